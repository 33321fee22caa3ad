"""Run ONE hot-path operator a few times (for ncu captures and quick timing).
    python tools/profile_case.py fcb35|fcb33|fcb53|fused35|bb128s2|bb128|bb256|bb512s2|headconv|corr|corrpairs|corrsweep|roialign [--frames 72] [--reps 5] [--backend auto]
"""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stmask_b200 import ops  # noqa: E402
from stmask_b200.hotpath import fpn_level_sizes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("case")
ap.add_argument("--frames", type=int, default=72)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--backend", default="auto")
ap.add_argument("--offset-scale", type=float, default=2.0)
ap.add_argument("--hint", type=int, default=0, help="STM_DCN_HINT_* bits (16 rows128, 32 rows256, 64 no-pair, 256 deep pipe)")
ap.add_argument("--nhwc-offsets", action="store_true", help="bb*: offsets / mask logits as a channels-last [B, Ho, Wo, 32] tensor (round-2 first half)")
a = ap.parse_args()
dev = "cuda"
torch.manual_seed(0)
F = a.frames


def timeit(fn, flops=None, nbytes=None):
    fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
    for s, e in evs:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ms = [s.elapsed_time(e) for s, e in evs]
    # back-to-back launches between ONE pair of events: launch latency is hidden by the queue
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = max(10, a.reps * 4)
    s0.record()
    for _ in range(nb):
        fn()
    e0.record()
    torch.cuda.synchronize()
    m = min(statistics.median(ms), s0.elapsed_time(e0) / nb)
    msg = f"{a.case}: {m:.4f} ms (single-launch median {statistics.median(ms):.4f}, back-to-back {s0.elapsed_time(e0) / nb:.4f})"
    if flops:
        msg += f"  {flops / m / 1e9:.1f} TFLOP/s"
    if nbytes:
        msg += f"  {nbytes / m / 1e6:.1f} GB/s"
    print(msg)


if a.case.startswith("fcb"):
    kh, kw = {"fcb33": (3, 3), "fcb35": (3, 5), "fcb53": (5, 3)}[a.case]
    spec = ops.ConvSpec(256, 256, (kh, kw), 1, ((kh - 1) // 2, (kw - 1) // 2))
    w = (torch.randn(256, 256, kh, kw, device=dev) / (256 * kh * kw) ** 0.5).bfloat16()
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    lv = fpn_level_sizes()
    xs = [torch.randn(F, 256, h, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for h, ww in lv]
    offs = [torch.randn(F, 2 * kh * kw, h, ww, device=dev) * a.offset_scale for h, ww in lv]
    outs = ops.deform_conv2d_multi(xs, offs, None, wp, None, spec, relu=True, backend=a.backend, hint=a.hint)
    px = sum(h * ww for h, ww in lv)
    print(ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, a.backend, a.hint))
    timeit(lambda: ops.deform_conv2d_multi(xs, offs, None, wp, None, spec, relu=True, backend=a.backend, outs=outs, hint=a.hint),
           flops=2.0 * F * px * 256 * 256 * kh * kw)
elif a.case.startswith("fused"):      # FCB with the offsets derived inside the kernel from the box deltas (what the step launches)
    kh, kw = {"fused33": (3, 3), "fused35": (3, 5), "fused53": (5, 3)}[a.case]
    spec = ops.ConvSpec(256, 256, (kh, kw), 1, ((kh - 1) // 2, (kw - 1) // 2))
    w = (torch.randn(256, 256, kh, kw, device=dev) / (256 * kh * kw) ** 0.5).bfloat16()
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    lv = fpn_level_sizes()
    xs = [torch.randn(F, 256, h, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for h, ww in lv]
    deltas = [torch.randn(F, 4, h, ww, device=dev) for h, ww in lv]
    w_off = torch.randn(2 * kh * kw, 4, 1, 1, device=dev) * 0.5
    outs = ops.deform_conv2d_fcb_multi(xs, deltas, wp, spec, w_off, relu=True, hint=a.hint)
    px = sum(h * ww for h, ww in lv)
    print(ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, a.backend, a.hint, fcb=True))
    timeit(lambda: ops.deform_conv2d_fcb_multi(xs, deltas, wp, spec, w_off, relu=True, outs=outs, hint=a.hint),
           flops=2.0 * F * px * 256 * 256 * kh * kw)
elif a.case.startswith("bb"):
    C, H, W, s = {"bb128s2": (128, 96, 160, 2), "bb128": (128, 48, 80, 1), "bb256s2": (256, 48, 80, 2), "bb256": (256, 24, 40, 1),
                  "bb512s2": (512, 24, 40, 2)}[a.case]
    spec = ops.ConvSpec(C, C, 3, s, 1)
    Ho, Wo = spec.out_hw(H, W)
    w = (torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5).bfloat16()
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    x = torch.randn(F, C, H, W, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    om = torch.randn(F, 32, Ho, Wo, device=dev) * (a.offset_scale / 2.0)   # fp32, plane-major, as the predictor writes it
    if a.nhwc_offsets:
        om = om.contiguous(memory_format=torch.channels_last)
    bias = torch.randn(C, device=dev)
    outs = ops.deform_conv2d_multi([x], [om[:, :18]], [om[:, 18:27]], wp, bias, spec, mask_sigmoid=True, backend=a.backend, hint=a.hint)
    print(ops.deform_conv2d_variant([tuple(x.shape)], spec, torch.bfloat16, a.backend, a.hint))
    timeit(lambda: ops.deform_conv2d_multi([x], [om[:, :18]], [om[:, 18:27]], wp, bias, spec, mask_sigmoid=True, backend=a.backend, outs=outs, hint=a.hint),
           flops=2.0 * F * Ho * Wo * C * C * 9)
    # its offset / mask-logit predictor: the plain-conv mode of the same main loop, fp32 output
    pc = ops.PlainConv()
    cw = (torch.randn(27, C, 3, 3, device=dev) * 0.02).bfloat16()
    cb = torch.randn(27, device=dev).bfloat16()
    pc([x], cw, cb, s, 1, 1, out_f32=True, out_planar=True)
    a.case += ".predictor"
    timeit(lambda: pc([x], cw, cb, s, 1, 1, out_f32=True, out_planar=True), flops=2.0 * F * Ho * Wo * C * 32 * 9)
elif a.case == "headconv":      # a 256 -> 256 3x3 conv + ReLU of the prediction head over P3..P7: ONE launch of the TMA shifted-view kernel
    lv = fpn_level_sizes()
    spec = ops.ConvSpec(256, 256, 3, 1, 1)
    w = (torch.randn(256, 256, 3, 3, device=dev) / (256 * 9) ** 0.5).bfloat16()
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    bias = torch.randn(256, device=dev)
    xs = [torch.randn(F, 256, h, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for h, ww in lv]
    outs = ops.deform_conv2d_multi(xs, [None] * 5, None, wp, bias, spec, relu=True, hint=a.hint)
    print(ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, a.backend, a.hint, zero_offset=True))
    timeit(lambda: ops.deform_conv2d_multi(xs, [None] * 5, None, wp, bias, spec, relu=True, outs=outs, hint=a.hint),
           flops=2.0 * F * sum(h * ww for h, ww in lv) * 256 * 256 * 9)
elif a.case == "corr":
    n = F - 1
    x1 = torch.randn(n, 256, 24, 40, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    x2 = torch.randn_like(x1)
    t1, t2 = torch.randn_like(x1), torch.randn_like(x1)
    foff = None if os.environ.get("STM_CORR_UNPADDED") else 128
    fn = lambda: ops.correlation(x1, x2, 11, 1, scale=1 / 256, relu=True, feats=(t1, t2), channels_last=True, backend=a.backend,
                                 feat_channel_offset=foff)
    print("algorithmic (SURVEY 8d) GB/s = reported x", (633 + 512) / (633 + 1024.0))
    timeit(fn, nbytes=n * 960 * (633 + 2 * 256 + 2 * 256) * 2.0)
elif a.case == "corrpairs":      # the hot path's temporal fusion: (t-1, t) pairs read in place from one F-frame clip
    from stmask_b200 import sharding
    x = torch.randn(F, 256, 24, 40, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    t = torch.randn_like(x)
    ri, ni = sharding.pair_index_tensors(sharding.make_plan(1, F, 1), 0, x.device)
    fn = lambda: ops.correlation_pairs(x, ri, ni, 11, 1, scale=1 / 256, relu=True, feats=t, feat_channel_offset=128)
    n = F - 1
    print("algorithmic (SURVEY 8d) GB/s = reported x", (633 + 512) / (633 + 1024.0))
    timeit(fn, nbytes=n * 960 * (633 + 2 * 256 + 2 * 256) * 2.0)
elif a.case == "roialign":       # bbox_feat_extractor on the padded concat buffer: 100 boxes per frame pair, 7x7
    n = F - 1
    x = torch.randn(n, 640, 24, 40, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    nb = 100
    g = torch.Generator(device=dev).manual_seed(0)
    xy = torch.rand(n * nb, 2, device=dev, generator=g) * torch.tensor([30.0, 16.0], device=dev)
    wh = 2 + torch.rand(n * nb, 2, device=dev, generator=g) * torch.tensor([10.0, 8.0], device=dev)
    rois = torch.cat([torch.arange(n, device=dev).repeat_interleave(nb)[:, None].float(), xy, xy + wh], 1)
    out = ops.roi_align(x, rois, 7)
    # bytes: every output element written once + (per sample point) 4 corner vectors read; report output-side GB/s
    timeit(lambda: ops.roi_align(x, rois, 7), nbytes=float(out.numel() * 2))
elif a.case == "corrsweep":       # BASELINE.json configs[1]: batch 8 over P3..P7, plain cost volume
    lv = fpn_level_sizes()
    xs = [(torch.randn(8, 256, h, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last),
           torch.randn(8, 256, h, ww, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)) for h, ww in lv]
    px = sum(h * ww for h, ww in lv)
    fn = lambda: ops.correlation_multi([p for p, _ in xs], [q for _, q in xs], 11, 1, scale=1 / 256, leaky_slope=0.1)   # ONE grouped launch
    timeit(fn, nbytes=8 * px * (2 * 256 + 121) * 2.0)
else:
    raise SystemExit("unknown case")
