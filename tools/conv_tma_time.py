"""Time the plain-conv launches of the hot path (DCN offset/mask predictors at the bench's 1024 frames, head convs) on the
TMA shifted-view kernel and on the gather main loop (STM_DCN_HINT_GATHER).  CUDA events, L2-sized inputs."""
import sys
import torch
from stmask_b200 import ops, _lib as L

dev = torch.device("cuda:0")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
EXP = int(sys.argv[2]) if len(sys.argv) > 2 else 0
LAYERS = [("pred C128 96x160 s2", 128, 32, 3, 2, 1, [(F, 96, 160)], True),
          ("pred C128 48x80 s1", 128, 32, 3, 1, 1, [(F, 48, 80)], True),
          ("pred C256 48x80 s2", 256, 32, 3, 2, 1, [(F, 48, 80)], True),
          ("pred C256 24x40 s1", 256, 32, 3, 1, 1, [(F, 24, 40)], True),
          ("pred C512 24x40 s2", 512, 32, 3, 2, 1, [(F, 24, 40)], True),
          ("head 256->256 3x3 P3..P7", 256, 256, 3, 1, 1, [(F // 8, 48, 80), (F // 8, 24, 40), (F // 8, 12, 20), (F // 8, 6, 10), (F // 8, 3, 5)], False),
          ("head 256->1024 3x3 P3..P7", 256, 1024, 3, 1, 1, [(F // 8, 48, 80), (F // 8, 24, 40), (F // 8, 12, 20), (F // 8, 6, 10), (F // 8, 3, 5)], False)]
for name, cin, cout, k, s, pad, maps, f32 in LAYERS:
    spec = ops.ConvSpec(cin, cout, k, s, pad)
    xs = [torch.randn((b, h, w, cin), device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2) for b, h, w in maps]
    w = (torch.randn((cout, cin, k, k), device=dev) / (cin * k * k) ** 0.5).bfloat16()
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    bias = torch.randn(cout, device=dev)
    outs = None
    res = {}
    for label, hint in (("tma", 0), ("gather", EXP << 20)):
        v = ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, zero_offset=True, hint=hint)
        ys = ops.deform_conv2d_multi(xs, [None] * len(xs), None, wp, bias, spec, out_f32=f32, hint=hint)
        for _ in range(2):
            ops.deform_conv2d_multi(xs, [None] * len(xs), None, wp, bias, spec, out_f32=f32, hint=hint, outs=ys)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            ops.deform_conv2d_multi(xs, [None] * len(xs), None, wp, bias, spec, out_f32=f32, hint=hint, outs=ys)
        e1.record()
        torch.cuda.synchronize()
        res[label] = (e0.elapsed_time(e1) / 5, v, ys)
    rows = sum(b * ((h + 2 * pad - k) // s + 1) * ((w_ + 2 * pad - k) // s + 1) for b, h, w_ in maps)
    flops = 2.0 * rows * cout * cin * k * k
    in_bytes = sum(x.numel() * 2 for x in xs)
    d = max(float((a.float() - b.float()).abs().max()) for a, b in zip(res["tma"][2], res["gather"][2]))
    print(f"{name}: tma {res['tma'][0]:.3f} ms ({flops / res['tma'][0] / 1e9:.0f} TF/s, input {in_bytes / res['tma'][0] / 1e6:.0f} GB/s)  "
          f"variant (exp bits) {res['gather'][0]:.3f} ms  max|diff| {d:.3g}\n    {res['tma'][1]}\n    {res['gather'][1]}", flush=True)
    del xs, res
