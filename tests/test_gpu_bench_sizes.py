"""Parity at the sizes the benchmark actually runs ("parity hole").

Every tcgen05 instantiation that produces a bench number is compared with the CPU oracle on DEFORMED
samples (offsets ~ N(0, 2^2) px, masks, bias, stride 2, deform_groups 4) at a problem size that makes the
launcher pick it, and the test asserts WHICH instantiation ran through `stm_deform_conv2d_variant`
(`ops.deform_conv2d_variant`).  The scheduling hints force the other CTA shapes through the same
entry point, so 128-row / 256-row and single-CTA / CTA-pair kernels all see the same inputs.

bf16 tolerance 1e-2 (north star); inputs are rounded to bf16 first and the oracle runs on the rounded values.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16
TOL = 1e-2
FPN = [(48, 80), (24, 40), (12, 20), (6, 10), (3, 5)]


def _ops():
    from stmask_b200 import _lib, ops
    return ops, _lib


def q(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(BF16).float().numpy()


def dev(a, dtype, device, cl=True):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device=device, dtype=dtype)
    return t.contiguous(memory_format=torch.channels_last) if cl and t.dim() == 4 else t


def _variants_to_check(ops, L, shapes, spec):
    """(hint, expected substrings) for every CTA shape the launcher offers for this problem."""
    auto = ops.deform_conv2d_variant(shapes, spec, BF16)
    out = [(0, auto)]
    for hint in (L.DCN_HINT_NO_PAIR, L.DCN_HINT_ROWS128, L.DCN_HINT_ROWS128 | L.DCN_HINT_NO_PAIR, L.DCN_HINT_ROWS256,
                 L.DCN_HINT_ROWS256 | L.DCN_HINT_NO_PAIR, L.DCN_HINT_DEEP_PIPE, L.DCN_HINT_TWO_CTAS,
                 L.DCN_HINT_TWO_CTAS | L.DCN_HINT_NO_PAIR):
        v = ops.deform_conv2d_variant(shapes, spec, BF16, hint=hint)
        if all(v != o[1] for o in out):
            out.append((hint, v))
    return out


# ------------------------------------------------------------------------------------------
# BASELINE.json configs[1]: operator sweep, batch 8, P3..P7 in ONE grouped launch, dg 1 and 4
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dg", [1, 4])
@pytest.mark.parametrize("kernel", [(3, 3), (3, 5), (5, 3)], ids=lambda k: f"{k[0]}x{k[1]}")
def test_fcb_sweep_grouped_launch_vs_oracle(cuda_device, kernel, dg):
    ops, L = _ops()
    kh, kw = kernel
    pad = ((kh - 1) // 2, (kw - 1) // 2)
    F = 8
    rng = np.random.default_rng(kh * 100 + kw * 10 + dg)
    spec = ops.ConvSpec(256, 256, kernel, 1, pad, 1, 1, dg)
    w = q(rng.standard_normal((256, 256, kh, kw)) / np.sqrt(256 * kh * kw))
    xs = [q(rng.standard_normal((F, 256, h, ww))) for h, ww in FPN]
    offs = [(rng.standard_normal((F, dg * 2 * kh * kw, h, ww)) * 2.0).astype(np.float32) for h, ww in FPN]
    wants = [np.maximum(oracle.deform_conv2d(x, o, w, padding=pad, deform_groups=dg), 0) for x, o in zip(xs, offs)]
    wp = ops.pack_weight(dev(w, BF16, cuda_device, cl=False), spec, BF16)
    xd = [dev(x, BF16, cuda_device) for x in xs]
    od = [dev(o, torch.float32, cuda_device, cl=False) for o in offs]
    shapes = [tuple(x.shape) for x in xs]
    checked = _variants_to_check(ops, L, shapes, spec)
    # 40 920 rows: the bench instantiation (two 128-row CTAs per SM, each half of a cta_group::2 pair)
    assert "tcgen05 rows=128 n=256 pair=1" in checked[0][1] and "ctas_per_sm=2" in checked[0][1], checked[0][1]
    assert any("rows=256" in v and "pair=0" in v for _, v in checked) and any("rows=256" in v and "pair=1" in v for _, v in checked), checked
    assert len(checked) >= 2, checked
    for hint, variant in checked:
        ys = ops.deform_conv2d_multi(xd, od, None, wp, None, spec, relu=True, hint=hint)
        torch.cuda.synchronize()
        for lvl, (y, want) in enumerate(zip(ys, wants)):
            err = rel_err(y.float().cpu().numpy(), want)
            assert err <= TOL, (variant, lvl, err)


# ------------------------------------------------------------------------------------------
# backbone DCNv2 layers at sizes that select the benchmarked instantiations
# ------------------------------------------------------------------------------------------
BACKBONE = [
    # C,  H,  W, stride, frames, expected in variant
    (256, 48, 80, 2, 20, "rows=256 n=256 pair=0"),     # layer3 block 0 (s2): 19 200 rows >= 18 944
    (256, 24, 40, 1, 20, "rows=256 n=256 pair=0"),     # layer3
    (256, 24, 40, 1, 40, "rows=128 n=256 pair=1"),     # layer3 at >= 37 888 rows (the 72-frame bench): paired two-CTA shape
    (256, 48, 80, 2, 40, "rows=128 n=256 pair=1"),
    (512, 24, 40, 2, 40, "rows=256 n=256 pair=0"),     # layer4 block 0: two N tiles, 9 600 rows >= 9 472
    (512, 12, 20, 1, 40, "rows=256 n=256 pair=0"),
    (512, 12, 20, 1, 80, "rows=128 n=256 pair=1"),     # layer4 at 19 200 rows x 2 N tiles (c5 workload sizes)
    (128, 48, 80, 1, 8, "rows=128 n=128"),       # layer2: two 8-warp CTAs per SM
    (128, 96, 160, 2, 6, "rows=128 n=128"),      # layer2 block 0 (s2)
]


@pytest.mark.parametrize("case", BACKBONE, ids=lambda c: f"C{c[0]}_{c[1]}x{c[2]}_s{c[3]}_F{c[4]}")
def test_backbone_dcnv2_bench_sizes_vs_oracle(cuda_device, case):
    ops, L = _ops()
    C, H, W, s, F, expect = case
    rng = np.random.default_rng(C + H + s)
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    spec = ops.ConvSpec(C, C, 3, s, 1)
    x = q(rng.standard_normal((F, C, H, W)))
    w = q(rng.standard_normal((C, C, 3, 3)) / np.sqrt(9 * C))
    bias = rng.standard_normal(C).astype(np.float32)
    off = (rng.standard_normal((F, 18, Ho, Wo)) * 2.0).astype(np.float32)
    logits = rng.standard_normal((F, 9, Ho, Wo)).astype(np.float32)
    mask = (1.0 / (1.0 + np.exp(-logits.astype(np.float64)))).astype(np.float32)
    want = oracle.deform_conv2d(x, off, w, bias, mask, stride=s, padding=1)
    # offsets and mask logits as ONE [F, Ho, Wo, 32] fp32 tensor (what the offset predictor produces): channel views
    om = np.zeros((F, 32, Ho, Wo), np.float32)
    om[:, :18], om[:, 18:27] = off, logits
    omd = dev(om, torch.float32, cuda_device)
    wp = ops.pack_weight(dev(w, BF16, cuda_device, cl=False), spec, BF16)
    xd = dev(x, BF16, cuda_device)
    bd = dev(bias, torch.float32, cuda_device)
    checked = _variants_to_check(ops, L, [tuple(x.shape)], spec)
    assert expect in checked[0][1], checked[0][1]
    for hint, variant in checked:
        y = ops.deform_conv2d_multi([xd], [omd[:, :18]], [omd[:, 18:27]], wp, bd, spec, mask_sigmoid=True, hint=hint)[0]
        torch.cuda.synchronize()
        err = rel_err(y.float().cpu().numpy(), want)
        assert err <= TOL, (variant, err)


def test_corner_weight_rounding_bound_with_large_magnitude_features(cuda_device):
    """The tcgen05 producer rounds the four corner weights (bilinear x mask) to bf16 before the fp32 blend
    (relative error <= 2^-9 per weight, on top of the bf16 rounding of the blended A element that any bf16
    GEMM has).  Large-magnitude features with a large mean make the weight error visible if it were not bounded:
    the result must stay within the bf16 tolerance of the oracle."""
    ops, L = _ops()
    rng = np.random.default_rng(99)
    F, C, H, W = 8, 256, 48, 80
    x = q(rng.standard_normal((F, C, H, W)) * 50.0 + 100.0)
    w = q(rng.standard_normal((C, C, 3, 3)) / np.sqrt(9 * C))
    off = (rng.standard_normal((F, 18, H, W)) * 2.0).astype(np.float32)
    mask = rng.random((F, 9, H, W)).astype(np.float32)
    want = oracle.deform_conv2d(x, off, w, None, mask, padding=1)
    y = ops.deform_conv2d(dev(x, BF16, cuda_device), dev(off, torch.float32, cuda_device, False), dev(w, BF16, cuda_device, False),
                          None, dev(mask, torch.float32, cuda_device, False), padding=1)
    assert rel_err(y.float().cpu().numpy(), want) <= TOL


# ------------------------------------------------------------------------------------------
# correlation sweep: batch 8, P3..P7, dilation_patch 1 and 2
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [1, 2])
def test_correlation_sweep_vs_oracle(cuda_device, d):
    ops, _ = _ops()
    rng = np.random.default_rng(40 + d)
    for h, w in FPN:
        x1, x2 = q(rng.standard_normal((8, 256, h, w))), q(rng.standard_normal((8, 256, h, w)))
        want = oracle.correlation(x1, x2, 11, d).reshape(8, 121, h, w)
        for cl in (False, True):
            got = ops.correlation(dev(x1, BF16, cuda_device), dev(x2, BF16, cuda_device), 11, d, channels_last=cl,
                                  out_dtype=torch.float32)
            assert got.shape == (8, 121, h, w)
            assert rel_err(got.cpu().numpy(), want) <= TOL, (h, w, d, cl)


@pytest.mark.parametrize("d", [1, 2])
@pytest.mark.parametrize("padded", [True, False])
def test_correlation_grouped_multi_level_launch_vs_oracle(cuda_device, d, padded):
    """BASELINE.json configs[1]: the cost volumes of P3..P7 at batch 8 in ONE launch, channels-last."""
    ops, L = _ops()
    rng = np.random.default_rng(50 + d)
    x1s = [q(rng.standard_normal((8, 256, h, w))) for h, w in FPN]
    x2s = [q(rng.standard_normal((8, 256, h, w))) for h, w in FPN]
    n0 = L.launch_count()
    outs = ops.correlation_multi([dev(a, BF16, cuda_device) for a in x1s], [dev(b, BF16, cuda_device) for b in x2s], 11, d,
                                 scale=1.0 / 256, leaky_slope=0.1, padded=padded)
    torch.cuda.synchronize()
    assert L.launch_count() - n0 == 1
    for (h, w), a, b, o in zip(FPN, x1s, x2s, outs):
        assert o.shape == (8, 128 if padded else 121, h, w) and o.stride(1) == 1
        want = oracle.correlate(a, b, 11, d)
        assert rel_err(o[:, :121].float().cpu().numpy(), want) <= TOL, (h, w, d)
        if padded:
            assert float(o[:, 121:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------
# FCB with the offsets derived INSIDE the sampling kernel from the box deltas
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,dg", [("ada", 1), ("ada", 4), ("ali", 1)])
@pytest.mark.parametrize("kernel", [(3, 3), (3, 5), (5, 3)], ids=lambda k: f"{k[0]}x{k[1]}")
def test_fcb_fused_offsets_vs_oracle(cuda_device, kernel, mode, dg):
    """stm_deform_conv2d_fcb_fwd over P3..P7 at batch 8 (one launch, no offset tensors) against the oracle's
    FeatureAlign restatement (offsets from the deltas in fp64, then the deformable conv + ReLU); also every CTA shape."""
    ops, L = _ops()
    kh, kw = kernel
    pad = ((kh - 1) // 2, (kw - 1) // 2)
    rng = np.random.default_rng(kh * 7 + kw * 3 + dg + (1 if mode == "ada" else 0))
    spec = ops.ConvSpec(256, 256, kernel, 1, pad, 1, 1, dg)
    w = q(rng.standard_normal((256, 256, kh, kw)) / np.sqrt(256 * kh * kw))
    xs = [q(rng.standard_normal((8, 256, h, ww))) for h, ww in FPN]
    deltas = [rng.standard_normal((8, 4, h, ww)).astype(np.float32) for h, ww in FPN]
    w_off = (rng.standard_normal((dg * 2 * kh * kw, 4, 1, 1)) * 0.5).astype(np.float32) if mode == "ada" else None
    wants = [oracle.feature_align(x, dl, w, kernel, w_offset=w_off, deform_groups=dg)[0] for x, dl in zip(xs, deltas)]
    wp = ops.pack_weight(dev(w, BF16, cuda_device, cl=False), spec, BF16)
    xd = [dev(x, BF16, cuda_device) for x in xs]
    dd = [dev(dl, torch.float32, cuda_device, cl=i % 2 == 0) for i, dl in enumerate(deltas)]      # NHWC and NCHW delta tensors
    wod = dev(w_off, torch.float32, cuda_device, cl=False) if w_off is not None else None
    n0 = L.launch_count()
    for hint in (0, L.DCN_HINT_NO_PAIR, L.DCN_HINT_ROWS128 | L.DCN_HINT_NO_PAIR, L.DCN_HINT_ROWS256):
        ys = ops.deform_conv2d_fcb_multi(xd, dd, wp, spec, wod, relu=True, hint=hint)
        torch.cuda.synchronize()
        for lvl, (y, want) in enumerate(zip(ys, wants)):
            err = rel_err(y.float().cpu().numpy(), want)
            assert err <= TOL, (mode, hint, lvl, err)
    assert L.launch_count() - n0 == 4                               # one kernel per call: no offset kernels


def test_feature_align_module_uses_the_fused_path(cuda_device):
    from stmask_b200.feature_align import FeatureAlign
    ops, L = _ops()
    rng = np.random.default_rng(3)
    for mode in ("ada", "ali"):
        m = FeatureAlign(256, 41, (3, 5), deformable_groups=1, use_pred_offset=mode == "ada").to(cuda_device)
        torch.nn.init.normal_(m.conv_adaption.weight, std=0.02)
        if mode == "ada":
            torch.nn.init.normal_(m.conv_offset.weight, std=0.5)
        m.conv_adaption.to(BF16)
        xs = [dev(q(rng.standard_normal((2, 256, h, w))), BF16, cuda_device) for h, w in FPN[1:4]]
        bs = [dev(rng.standard_normal((2, 4, h, w)), torch.float32, cuda_device, cl=False) for h, w in FPN[1:4]]
        with torch.no_grad():
            m.calibrate_levels(xs, bs)                            # packs the weight once
            n0 = L.launch_count()
            fused = m.calibrate_levels(xs, bs)
        assert m._fused is True and L.launch_count() - n0 == 1     # ONE kernel: no offset kernels, no offset tensors
        with torch.no_grad():
            two_step = ops.deform_conv2d_multi(xs, [m.offsets(b) for b in bs], None,
                                               m.conv_adaption._cache.weight(m.conv_adaption.weight, m.conv_adaption.spec(), BF16), None,
                                               m.conv_adaption.spec(), relu=True)
        for a, b in zip(fused, two_step):
            assert rel_err(a.float().cpu().numpy(), b.float().cpu().numpy()) <= 1e-2     # same samples up to fp32 rounding of the offsets: bf16 ulps


# ------------------------------------------------------------------------------------------
# K-block order: chunk-major with resident sample records (default, dg == 1) vs tap-major (STM_DCN_HINT_TAP_MAJOR)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [("fcb", (3, 5), 256, 256, 1), ("fcb", (5, 3), 256, 256, 1), ("dcnv2", (3, 3), 128, 128, 1),
                                  ("dcnv2", (3, 3), 256, 256, 2), ("dcnv2", (3, 3), 512, 512, 2)],
                         ids=lambda c: f"{c[0]}_{c[1][0]}x{c[1][1]}_C{c[2]}_s{c[4]}")
def test_k_order_variants_vs_oracle(cuda_device, case):
    """Both K orders of the tcgen05 sampling kernel against the oracle, and the plan reports which one runs."""
    ops, L = _ops()
    kind, (kh, kw), cin, cout, s = case
    pad = ((kh - 1) // 2, (kw - 1) // 2)
    rng = np.random.default_rng(kh * 5 + kw + cin + s)
    B, H, W = 6, 26, 44
    spec = ops.ConvSpec(cin, cout, (kh, kw), s, pad)
    Ho, Wo = spec.out_hw(H, W)
    w = q(rng.standard_normal((cout, cin, kh, kw)) / np.sqrt(cin * kh * kw))
    x = q(rng.standard_normal((B, cin, H, W)))
    wp = ops.pack_weight(dev(w, BF16, cuda_device, cl=False), spec, BF16)
    xd = dev(x, BF16, cuda_device)
    v_auto = ops.deform_conv2d_variant([tuple(xd.shape)], spec, BF16, fcb=kind == "fcb")
    v_chunk = ops.deform_conv2d_variant([tuple(xd.shape)], spec, BF16, hint=L.DCN_HINT_CHUNK_MAJOR, fcb=kind == "fcb")
    v_tap = ops.deform_conv2d_variant([tuple(xd.shape)], spec, BF16, hint=L.DCN_HINT_TAP_MAJOR, fcb=kind == "fcb")
    assert "korder=chunk" in v_chunk and "korder=tap" in v_tap, (v_chunk, v_tap)
    assert "korder=chunk" in v_auto, v_auto
    if kind == "fcb":
        deltas = rng.standard_normal((B, 4, Ho, Wo)).astype(np.float32)
        w_off = (rng.standard_normal((2 * kh * kw, 4, 1, 1)) * 0.5).astype(np.float32)
        want = oracle.feature_align(x, deltas, w, (kh, kw), w_offset=w_off)[0]
        dd, wod = dev(deltas, torch.float32, cuda_device), dev(w_off, torch.float32, cuda_device, cl=False)
        run = lambda hint: ops.deform_conv2d_fcb_multi([xd], [dd], wp, spec, wod, relu=True, hint=hint)[0]
    else:
        off = (rng.standard_normal((B, 2 * kh * kw, Ho, Wo)) * 2).astype(np.float32)
        msk = rng.standard_normal((B, kh * kw, Ho, Wo)).astype(np.float32)
        bias = rng.standard_normal(cout).astype(np.float32)
        want = oracle.deform_conv2d(x, off, w, bias, 1.0 / (1.0 + np.exp(-msk)), stride=s, padding=pad)
        od, md, bd = dev(off, torch.float32, cuda_device), dev(msk, torch.float32, cuda_device), dev(bias, torch.float32, cuda_device, cl=False)
        run = lambda hint: ops.deform_conv2d_multi([xd], [od], [md], wp, bd, spec, mask_sigmoid=True, hint=hint)[0]
    CM = L.DCN_HINT_CHUNK_MAJOR
    for hint in (0, CM, L.DCN_HINT_TAP_MAJOR, L.DCN_HINT_ROWS256 | CM, L.DCN_HINT_ROWS128 | L.DCN_HINT_NO_PAIR | CM, L.DCN_HINT_TWO_CTAS | CM):
        y = run(hint)
        torch.cuda.synchronize()
        err = rel_err(y.float().cpu().numpy(), want)
        assert err <= TOL, (case, hint, err)


@pytest.mark.parametrize("case", [("fcb", (3, 5), 256, 1, [(3, 24, 40), (2, 12, 20), (2, 8, 8)]), ("dcnv2", (3, 3), 128, 2, [(5, 48, 80)]),
                                  ("dcnv2", (3, 3), 256, 1, [(7, 24, 40)])], ids=lambda c: f"{c[0]}_C{c[2]}_s{c[3]}")
def test_patch_row_order_is_bit_identical_to_raster_order(cuda_device, case):
    """Maps whose sides are multiples of 8 enumerate their GEMM rows in 8x8 pixel patches (L1 / L2 locality of the gather);
    every output pixel's arithmetic is unchanged, so the result must equal the raster-order launch (STM_DCN_HINT_RASTER) bit for
    bit — and the oracle within tolerance."""
    ops, L = _ops()
    kind, (kh, kw), c, s, maps = case
    pad = ((kh - 1) // 2, (kw - 1) // 2)
    rng = np.random.default_rng(kh + kw + c + s)
    spec = ops.ConvSpec(c, c, (kh, kw), s, pad)
    w = q(rng.standard_normal((c, c, kh, kw)) / np.sqrt(c * kh * kw))
    wp = ops.pack_weight(dev(w, BF16, cuda_device, cl=False), spec, BF16)
    xs = [q(rng.standard_normal((b, c, h, ww))) for b, h, ww in maps]
    xd = [dev(x, BF16, cuda_device) for x in xs]
    outs_hw = [spec.out_hw(h, ww) for _, h, ww in maps]
    if kind == "fcb":
        deltas = [rng.standard_normal((b, 4, ho, wo)).astype(np.float32) for (b, _, _), (ho, wo) in zip(maps, outs_hw)]
        w_off = (rng.standard_normal((2 * kh * kw, 4, 1, 1)) * 0.5).astype(np.float32)
        wants = [oracle.feature_align(x, dl, w, (kh, kw), w_offset=w_off)[0] for x, dl in zip(xs, deltas)]
        dd, wod = [dev(dl, torch.float32, cuda_device) for dl in deltas], dev(w_off, torch.float32, cuda_device, cl=False)
        run = lambda hint: ops.deform_conv2d_fcb_multi(xd, dd, wp, spec, wod, relu=True, hint=hint)
    else:
        offs = [(rng.standard_normal((b, 2 * kh * kw, ho, wo)) * 2).astype(np.float32) for (b, _, _), (ho, wo) in zip(maps, outs_hw)]
        msks = [rng.standard_normal((b, kh * kw, ho, wo)).astype(np.float32) for (b, _, _), (ho, wo) in zip(maps, outs_hw)]
        wants = [oracle.deform_conv2d(x, o, w, None, 1.0 / (1.0 + np.exp(-m)), stride=s, padding=pad) for x, o, m in zip(xs, offs, msks)]
        od = [dev(o, torch.float32, cuda_device, cl=False) for o in offs]            # plane-major, as the DCN module's predictor writes them
        md = [dev(m, torch.float32, cuda_device, cl=False) for m in msks]
        run = lambda hint: ops.deform_conv2d_multi(xd, od, md, wp, None, spec, mask_sigmoid=True, hint=hint)
    for base in (0, L.DCN_HINT_ROWS256, L.DCN_HINT_TAP_MAJOR):
        ya, yb = run(base), run(base | L.DCN_HINT_RASTER)
        torch.cuda.synchronize()
        for a, b, want in zip(ya, yb, wants):
            assert torch.equal(a, b)
            assert rel_err(a.float().cpu().numpy(), want) <= TOL
