"""Determinism stress of the TMA shifted-view convolution: the same launch repeated N times must give bit-identical outputs
(races between the TMA ring, the MMA thread, the double-buffered TMEM accumulators and the fused-tap exchange buffer would show
up as flips).  python tools/conv_tma_stress.py [reps]"""
import sys
import torch
from stmask_b200 import ops

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
CASES = [("fused predictor C256 24x40", 256, 32, 1, [(256, 24, 40)], True, True), ("fused predictor C128 48x80", 128, 32, 1, [(96, 48, 80)], True, True),
         ("stride-2 predictor C128 96x160", 128, 32, 2, [(48, 96, 160)], True, True), ("stride-2 predictor C512 24x40 (weight ring)", 512, 32, 2, [(128, 24, 40)], True, False),
         ("head conv 256->256 P3..P7", 256, 256, 1, [(16, 48, 80), (16, 24, 40), (16, 12, 20), (16, 6, 10), (16, 3, 5)], False, False)]
for name, cin, cout, s, maps, f32, planar in CASES:
    g = torch.Generator(device=dev).manual_seed(cin + cout)
    spec = ops.ConvSpec(cin, cout, 3, s, 1)
    xs = [torch.randn((b, h, w, cin), generator=g, device=dev).bfloat16().permute(0, 3, 1, 2) for b, h, w in maps]
    w = (torch.randn((cout, cin, 3, 3), generator=g, device=dev) / (cin * 9) ** 0.5).bfloat16()
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    bias = torch.randn(cout, generator=g, device=dev)
    ref = ops.deform_conv2d_multi(xs, [None] * len(xs), None, wp, bias, spec, relu=True, out_f32=f32, out_planar=planar)
    ref = [r.clone() for r in ref]
    bad = 0
    for i in range(reps):
        ys = ops.deform_conv2d_multi(xs, [None] * len(xs), None, wp, bias, spec, relu=True, out_f32=f32, out_planar=planar)
        bad += sum(0 if torch.equal(a, b) else 1 for a, b in zip(ys, ref))
    torch.cuda.synchronize()
    print(f"{name}: {reps} launches, {bad} outputs differ from the first launch  [{ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, zero_offset=True)}]")
    assert bad == 0
print("conv_tma stress OK")
