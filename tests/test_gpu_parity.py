"""GPU parity tests: the CUDA path (through the C ABI, via the reference-facing Python API)
against the CPU oracle and the committed golden vectors.

Tolerances are the north star's: <= 1e-4 relative in fp32, <= 1e-2 in bf16, where relative
error is max|got - want| / max|want| over a feature tensor (conftest.rel_err).  For bf16 the
inputs are rounded to bf16 first and the oracle runs on the rounded values, so the figure
measures the kernel's arithmetic, not the input quantisation.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import golden_meta, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}
DTYPES = [torch.float32, torch.bfloat16]


def _ops():
    from stmask_b200 import ops
    return ops


def q(a, dtype):
    """numpy fp32 -> numpy fp32 rounded to `dtype`."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dtype).float().numpy()


def dev(a, dtype, device, channels_last=False):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device=device, dtype=dtype)
    if channels_last and t.dim() == 4:
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def run_dcn(device, dtype, x, offset, weight, bias=None, mask=None, backend="auto", channels_last=True,
            offset_dtype=torch.float32, **kw):
    ops = _ops()
    y = ops.deform_conv2d(dev(x, dtype, device, channels_last), dev(offset, offset_dtype, device) if offset is not None else None,
                          dev(weight, dtype, device), dev(bias, torch.float32, device) if bias is not None else None,
                          dev(mask, offset_dtype, device) if mask is not None else None, backend=backend, **kw)
    torch.cuda.synchronize()
    return y.float().cpu().numpy()


# ------------------------------------------------------------------------------------------
# deformable conv vs golden vectors (torchvision CPU) and vs the oracle
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_dcn_golden_torchvision(cuda_device, dtype):
    z = load_golden("dcn_torchvision.npz")
    for name, m in golden_meta(z).items():
        x, w = q(z[f"{name}.x"], dtype), q(z[f"{name}.weight"], dtype)
        bias = z[f"{name}.bias"] if f"{name}.bias" in z else None
        mask = z[f"{name}.mask"] if f"{name}.mask" in z else None
        kw = dict(stride=m["stride"], padding=m["padding"], dilation=m["dilation"], groups=m["groups"],
                  deform_groups=m["deform_groups"])
        got = run_dcn(cuda_device, dtype, x, z[f"{name}.offset"], w, bias, mask, **kw)
        want = z[f"{name}.y"] if dtype == torch.float32 else oracle.deform_conv2d(x, z[f"{name}.offset"], w, bias, mask, **kw)
        assert got.shape == want.shape, name
        assert rel_err(got, want) <= TOL[dtype], (name, rel_err(got, want))


def test_dcn_border_rule(cuda_device):
    z = load_golden("dcn_border.npz")
    x = z["x"]
    H, W = x.shape[2:]
    w = np.ones((1, 1, 1, 1), np.float32)
    for (h, wv), want in zip(z["positions"], z["values"]):
        off = np.zeros((1, 2, H, W), np.float32)
        off[0, 0, 0, 0], off[0, 1, 0, 0] = h, wv
        got = run_dcn(cuda_device, torch.float32, x, off, w)[0, 0, 0, 0]
        assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), (h, wv, got, want)


CASES = [
    # B, Cin, Cout, H, W, kh, kw, stride, dg, mask, bias
    (2, 256, 256, 12, 20, 3, 3, 1, 1, False, False),   # FCB 3x3 on P5
    (2, 256, 256, 12, 20, 3, 5, 1, 1, False, False),   # FCB 3x5
    (2, 256, 256, 12, 20, 5, 3, 1, 1, False, False),   # FCB 5x3
    (1, 256, 256, 3, 5, 5, 3, 1, 1, False, False),     # 5x3 on P7 (smaller than the kernel)
    (2, 256, 256, 6, 10, 3, 3, 1, 4, False, False),    # deform_groups = 4
    (1, 256, 256, 12, 20, 3, 5, 1, 4, False, False),
    (2, 128, 128, 24, 40, 3, 3, 2, 1, True, True),     # backbone layer2 block 0 style (stride 2)
    (2, 128, 128, 12, 20, 3, 3, 1, 1, True, True),
    (1, 512, 512, 6, 10, 3, 3, 1, 1, True, True),      # backbone layer4
    (1, 64, 96, 9, 11, 3, 3, 1, 1, True, False),
    (3, 256, 256, 23, 40, 3, 3, 1, 1, False, False),   # odd height (unpadded 360x640 level)
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("backend", ["simt", "auto"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_dcn_vs_oracle(cuda_device, dtype, backend, case):
    B, Cin, Cout, H, W, kh, kw, s, dg, use_mask, use_bias = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    pad = ((kh - 1) // 2, (kw - 1) // 2)
    Ho, Wo = oracle.out_size(H, kh, s, pad[0], 1), oracle.out_size(W, kw, s, pad[1], 1)
    x = q(rng.standard_normal((B, Cin, H, W)), dtype)
    w = q(rng.standard_normal((Cout, Cin, kh, kw)) / np.sqrt(Cin * kh * kw), dtype)
    off = (rng.standard_normal((B, dg * 2 * kh * kw, Ho, Wo)) * 2.0).astype(np.float32)
    mask = rng.random((B, dg * kh * kw, Ho, Wo)).astype(np.float32) if use_mask else None
    bias = rng.standard_normal(Cout).astype(np.float32) if use_bias else None
    kwargs = dict(stride=s, padding=pad, deform_groups=dg)
    want = oracle.deform_conv2d(x, off, w, bias, mask, **kwargs)
    got = run_dcn(cuda_device, dtype, x, off, w, bias, mask, backend=backend, **kwargs)
    assert rel_err(got, want) <= TOL[dtype], rel_err(got, want)


@pytest.mark.parametrize("dtype", DTYPES)
def test_dcn_layout_and_flag_variants(cuda_device, dtype):
    """NCHW-contiguous input, bf16 offsets, NHWC offsets, fused ReLU, sigmoid-on-logits: same numbers."""
    rng = np.random.default_rng(5)
    B, C, H, W = 2, 64, 10, 12
    x = q(rng.standard_normal((B, C, H, W)), dtype)
    w = q(rng.standard_normal((C, C, 3, 3)) / 24.0, dtype)
    off = q(rng.standard_normal((B, 18, H, W)) * 2.0, torch.bfloat16)      # exactly representable in bf16
    logits = q(rng.standard_normal((B, 9, H, W)), torch.bfloat16)
    mask = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
    want = oracle.deform_conv2d(x, off, w, None, mask.astype(np.float32), padding=1)
    ops = _ops()
    xd, wd = dev(x, dtype, cuda_device), dev(w, dtype, cuda_device)
    for off_dtype in (torch.float32, torch.bfloat16):
        for cl in (False, True):
            o = dev(off, off_dtype, cuda_device, cl)
            m = dev(logits, off_dtype, cuda_device, cl)
            y = ops.deform_conv2d(xd if not cl else xd.contiguous(memory_format=torch.channels_last), o, wd, None, m,
                                  padding=1, mask_sigmoid=True, relu=True)
            assert y.shape == (B, C, H, W)
            assert rel_err(y.float().cpu().numpy(), np.maximum(want, 0)) <= TOL[dtype]


def test_dcn_zero_offset_equals_cudnn_conv_full_size(cuda_device):
    """Size-independent property at the benchmark size: zero offsets, mask 1 == F.conv2d."""
    ops = _ops()
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False          # the cuDNN reference must be real fp32
    for (C, H, W, k, pad) in ((256, 48, 80, (3, 3), (1, 1)), (256, 48, 80, (3, 5), (1, 2)), (256, 24, 40, (5, 3), (2, 1))):
        x = torch.randn(8, C, H, W, device=cuda_device).contiguous(memory_format=torch.channels_last)
        w = torch.randn(C, C, *k, device=cuda_device) / (C * k[0] * k[1]) ** 0.5
        off = torch.zeros(8, 2 * k[0] * k[1], H, W, device=cuda_device)
        ref = torch.nn.functional.conv2d(x, w, padding=pad)
        got = ops.deform_conv2d(x, off, w, padding=pad)
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) <= 1e-4
        gb = ops.deform_conv2d(x.bfloat16(), off, w.bfloat16(), padding=pad)
        refb = torch.nn.functional.conv2d(x.bfloat16().float(), w.bfloat16().float(), padding=pad)
        assert rel_err(gb.float().cpu().numpy(), refb.cpu().numpy()) <= 1e-2
        # plain-conv mode of the same kernel (offset=None)
        gz = ops.deform_conv2d(x, None, w, padding=pad)
        assert rel_err(gz.cpu().numpy(), ref.cpu().numpy()) <= 1e-4


def test_dcn_linearity_full_size(cuda_device):
    """y(a*x1 + x2) == a*y(x1) + y(x2) with shared random offsets, P3 at batch 8 (fp32)."""
    ops = _ops()
    torch.manual_seed(1)
    x1 = torch.randn(8, 256, 48, 80, device=cuda_device).contiguous(memory_format=torch.channels_last)
    x2 = torch.randn_like(x1)
    w = torch.randn(256, 256, 3, 5, device=cuda_device) / (256 * 15) ** 0.5
    off = torch.randn(8, 30, 48, 80, device=cuda_device) * 2
    f = lambda x: ops.deform_conv2d(x, off, w, padding=(1, 2))
    lhs = f(2.5 * x1 + x2)
    rhs = 2.5 * f(x1) + f(x2)
    assert rel_err(lhs.cpu().numpy(), rhs.cpu().numpy()) <= 1e-4


def test_dcn_multi_level_launch_matches_single(cuda_device):
    ops = _ops()
    torch.manual_seed(2)
    spec = ops.ConvSpec(256, 256, (3, 5), 1, (1, 2))
    w = torch.randn(256, 256, 3, 5, device=cuda_device, dtype=torch.bfloat16) / 60
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    sizes = [(48, 80), (24, 40), (12, 20), (6, 10), (3, 5)]
    xs = [torch.randn(2, 256, h, ww, device=cuda_device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last) for h, ww in sizes]
    offs = [torch.randn(2, 30, h, ww, device=cuda_device) * 2 for h, ww in sizes]
    multi = ops.deform_conv2d_multi(xs, offs, None, wp, None, spec, relu=True)
    for x, o, ym in zip(xs, offs, multi):
        y1 = ops.deform_conv2d_multi([x], [o], None, wp, None, spec, relu=True)[0]
        assert torch.equal(y1, ym)
        want = np.maximum(oracle.deform_conv2d(x.float().cpu().numpy(), o.cpu().numpy(), w.float().cpu().numpy(), padding=(1, 2)), 0)
        assert rel_err(ym.float().cpu().numpy(), want) <= 1e-2


# ------------------------------------------------------------------------------------------
# module-level parity with the reference's call sites (golden fixtures)
# ------------------------------------------------------------------------------------------
def _oracle_dcn_module(x, weight, bias, com_w, com_b, stride, dtype):
    """The reference's DCN.forward chain on (rounded) numpy inputs: predictor conv -> chunk/cat -> sigmoid -> DCNv2."""
    k = com_w.shape[2]
    pad = (k - 1) // 2
    B, _, H, W = x.shape
    Ho, Wo = oracle.out_size(H, k, stride, pad, 1), oracle.out_size(W, k, stride, pad, 1)
    zero = np.zeros((B, 2 * k * k, Ho, Wo), np.float32)
    om = oracle.deform_conv2d(x, zero, com_w, com_b, None, stride=stride, padding=pad)      # zero offsets == regular conv
    n_off = 2 * k * k
    mask = 1.0 / (1.0 + np.exp(-om[:, n_off:].astype(np.float64)))
    return oracle.deform_conv2d(x, om[:, :n_off], weight, bias, mask.astype(np.float32), stride=stride, padding=pad), om


@pytest.mark.parametrize("dtype", DTYPES)
def test_backbone_dcn_module_vs_reference(cuda_device, dtype):
    """The drop-in `DCN` (own fp32-output offset/mask predictor + modulated deformable conv) against the golden
    from the reference's own Bottleneck DCN branch (backbone.py:20-26,45).  fp32: the golden itself, 1e-4.
    bf16: inputs and parameters rounded to bf16 first, the oracle runs the same chain on the rounded values, 1e-2
    (the predictor's accumulators stay fp32, so sampling positions are not rounded)."""
    from stmask_b200.backbone_dcn import make_bottleneck_dcn
    z = load_golden("backbone_dcn.npz")
    for name, stride in (("s1", 1), ("s2", 2)):
        m = make_bottleneck_dcn(16, stride=stride).to(cuda_device)
        with torch.no_grad():
            m.weight.copy_(torch.from_numpy(z[f"{name}.weight"]))
            m.bias.copy_(torch.from_numpy(z[f"{name}.bias"]))
            m.conv_offset_mask.weight.copy_(torch.from_numpy(z[f"{name}.com_w"]))
            m.conv_offset_mask.bias.copy_(torch.from_numpy(z[f"{name}.com_b"]))
        m = m.to(dtype)
        x = dev(z[f"{name}.dcn_x"], dtype, cuda_device)
        with torch.no_grad():
            y = m(x)
        if dtype == torch.float32:
            assert rel_err(y.float().cpu().numpy(), z[f"{name}.dcn_y"]) <= 1e-4
        want, _ = _oracle_dcn_module(q(z[f"{name}.dcn_x"], dtype), q(z[f"{name}.weight"], dtype), q(z[f"{name}.bias"], dtype),
                                     q(z[f"{name}.com_w"], dtype), q(z[f"{name}.com_b"], dtype), stride, dtype)
        assert rel_err(y.float().cpu().numpy(), want) <= TOL[dtype], (name, rel_err(y.float().cpu().numpy(), want))


@pytest.mark.parametrize("case", [(256, 24, 40, 1, 20), (128, 48, 80, 2, 6), (512, 12, 20, 1, 8)],
                         ids=lambda c: f"C{c[0]}_{c[1]}x{c[2]}_s{c[3]}_F{c[4]}")
def test_backbone_dcn_module_tcgen05_vs_oracle(cuda_device, case):
    """Same chain at backbone channel counts, bf16, so that BOTH kernels of the module are the tcgen05 ones (the
    plain-conv predictor with fp32 output and the sampling kernel); non-zero predictor weights."""
    from stmask_b200 import ops
    from stmask_b200.compat.dcn_v2 import DCN
    C, H, W, s, F = case
    rng = np.random.default_rng(C + s)
    x = q(rng.standard_normal((F, C, H, W)), torch.bfloat16)
    w = q(rng.standard_normal((C, C, 3, 3)) / np.sqrt(9 * C), torch.bfloat16)
    b = q(rng.standard_normal(C) * 0.1, torch.bfloat16)
    cw = q(rng.standard_normal((27, C, 3, 3)) * 0.05 / 3, torch.bfloat16)
    cb = q(rng.standard_normal(27) * 0.5, torch.bfloat16)
    m = DCN(C, C, 3, s, 1).to(cuda_device)
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(w)); m.bias.copy_(torch.from_numpy(b))
        m.conv_offset_mask.weight.copy_(torch.from_numpy(cw)); m.conv_offset_mask.bias.copy_(torch.from_numpy(cb))
    m = m.to(torch.bfloat16)
    xd = dev(x, torch.bfloat16, cuda_device, channels_last=True)
    pred_spec = ops.ConvSpec(C, 32, 3, s, 1)
    assert "plain=1" in ops.deform_conv2d_variant([tuple(xd.shape)], pred_spec, torch.bfloat16, zero_offset=True)
    with torch.no_grad():
        y = m(xd)
        om = m._predictor([xd], m.conv_offset_mask.weight, m.conv_offset_mask.bias, s, 1, 1, out_f32=True)[0]
    want, want_om = _oracle_dcn_module(x, w, b, cw, cb, s, torch.bfloat16)
    assert om.dtype == torch.float32 and om.shape[1] == 32
    assert rel_err(om[:, :27].cpu().numpy(), want_om) <= 1e-4          # fp32 accumulators stored as fp32
    assert float(om[:, 27:].abs().max()) == 0.0
    assert rel_err(y.float().cpu().numpy(), want) <= 1e-2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", ["ada", "ali"])
def test_feature_align_vs_reference(cuda_device, dtype, mode):
    from stmask_b200.feature_align import FeatureAlign
    z = load_golden("feature_align.npz")
    for name, ks in (("k3x3", (3, 3)), ("k3x5", (3, 5)), ("k5x3", (5, 3)), ("k5x3_p7", (5, 3))):
        k = f"{mode}.{name}"
        m = FeatureAlign(32, 41, kernel_size=ks, deformable_groups=1, use_pred_offset=(mode == "ada")).to(cuda_device)
        with torch.no_grad():
            m.conv_adaption.weight.copy_(torch.from_numpy(z[f"{k}.w_adaption"]))
            m.conv.weight.copy_(torch.from_numpy(z[f"{k}.w_conv"]))
            m.conv.bias.copy_(torch.from_numpy(z[f"{k}.b_conv"]))
            if mode == "ada":
                m.conv_offset.weight.copy_(torch.from_numpy(z[f"{k}.w_offset"]))
        x = dev(z[f"{k}.x"], dtype, cuda_device)
        shape = dev(z[f"{k}.shape"], torch.float32, cuda_device)
        if dtype == torch.bfloat16:
            m.conv_adaption.to(dtype)
            m.conv.to(dtype)
        with torch.no_grad():
            off = m.offsets(shape)
            dcn = m.calibrate_levels([x], [shape])[0]
            y = m(x, shape)
        assert rel_err(off.float().cpu().numpy(), z[f"{k}.offset"]) <= 1e-5, k
        if dtype == torch.float32:
            assert rel_err(dcn.cpu().numpy(), z[f"{k}.dcn_relu"]) <= 1e-4, k
            assert rel_err(y.cpu().numpy(), z[f"{k}.y"]) <= 1e-4, k
        else:
            want, _ = oracle.feature_align(q(z[f"{k}.x"], dtype), z[f"{k}.shape"], q(z[f"{k}.w_adaption"], dtype), ks,
                                           w_offset=z[f"{k}.w_offset"] if mode == "ada" else None)
            assert rel_err(dcn.float().cpu().numpy(), want) <= 1e-2, k


@torch.no_grad()
def test_drop_in_modules_match_functional(cuda_device):
    from dcn_v2 import DCN, DCNv2, dcn_v2_conv
    from mmcv.ops import DeformConv2d, ModulatedDeformConv2d, ModulatedDeformConv2dPack, deform_conv2d, modulated_deform_conv2d
    torch.manual_seed(3)
    x = torch.randn(2, 32, 9, 10, device=cuda_device)
    off = torch.randn(2, 18, 9, 10, device=cuda_device)
    mask = torch.rand(2, 9, 9, 10, device=cuda_device)
    m1 = DeformConv2d(32, 48, (3, 3), padding=(1, 1)).to(cuda_device)
    want = oracle.deform_conv2d(x.cpu().numpy(), off.cpu().numpy(), m1.weight.detach().cpu().numpy(), padding=1)
    assert rel_err(m1(x, off).detach().cpu().numpy(), want) <= 1e-4
    assert rel_err(deform_conv2d(x, off, m1.weight, 1, 1).detach().cpu().numpy(), want) <= 1e-4
    m2 = ModulatedDeformConv2d(32, 48, 3, padding=1).to(cuda_device)
    m2.bias.data.normal_()
    want2 = oracle.deform_conv2d(x.cpu().numpy(), off.cpu().numpy(), m2.weight.detach().cpu().numpy(),
                                 m2.bias.detach().cpu().numpy(), mask.cpu().numpy(), padding=1)
    assert rel_err(m2(x, off, mask).detach().cpu().numpy(), want2) <= 1e-4
    assert rel_err(modulated_deform_conv2d(x, off, mask, m2.weight, m2.bias, 1, 1).detach().cpu().numpy(), want2) <= 1e-4
    m3 = DCNv2(32, 48, 3, 1, 1).to(cuda_device)
    m3.load_state_dict(m2.state_dict())
    assert rel_err(m3(x, off, mask).detach().cpu().numpy(), want2) <= 1e-4
    assert rel_err(dcn_v2_conv(x, off, mask, m2.weight, m2.bias, 1, 1, 1, 1).detach().cpu().numpy(), want2) <= 1e-4
    # the *Pack / DCN modules: offsets from their own zero-initialised predictor => plain conv with mask 0.5
    for cls in (lambda: ModulatedDeformConv2dPack(32, 48, 3, padding=1), lambda: DCN(32, 48, 3, 1, 1)):
        m4 = cls().to(cuda_device)
        ref = 0.5 * torch.nn.functional.conv2d(x, m4.weight, None, padding=1) + m4.bias.view(1, -1, 1, 1)
        assert rel_err(m4(x).detach().cpu().numpy(), ref.detach().cpu().numpy()) <= 1e-4
    # weight update is picked up (packed-weight cache keyed on the version counter)
    with torch.no_grad():
        m1.weight.mul_(2.0)
    assert rel_err(m1(x, off).detach().cpu().numpy(), 2 * want) <= 1e-4


# ------------------------------------------------------------------------------------------
# correlation
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_correlate_golden_reference_call_site(cuda_device, dtype):
    from stmask_b200.temporal_fusion import correlate, correlate_concat
    z = load_golden("correlate.npz")
    for name in ("p11_d1", "p11_d2", "p11_p7", "p5_d1"):
        P, d = (int(v) for v in z[f"{name}.pd"])
        x1, x2 = q(z[f"{name}.x1"], dtype), q(z[f"{name}.x2"], dtype)
        got = correlate(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), P, d)
        want = z[f"{name}.y"] if dtype == torch.float32 else oracle.correlate(x1, x2, P, d)
        assert got.shape == want.shape
        assert rel_err(got.float().cpu().numpy(), want) <= TOL[dtype], name
    x1, x2 = q(z["p11_d1.x1"], dtype), q(z["p11_d1.x2"], dtype)
    ta, tb = q(z["concat.t_ref"], dtype), q(z["concat.t_next"], dtype)
    for cl in (True, False):
        got = correlate_concat(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), dev(ta, dtype, cuda_device),
                               dev(tb, dtype, cuda_device), channels_last=cl)
        want = np.maximum(np.concatenate([oracle.correlate(x1, x2, 11, 1), ta, tb], 1), 0)
        assert got.shape == want.shape
        assert rel_err(got.float().cpu().numpy(), want) <= TOL[dtype]
        if dtype == torch.float32:
            assert rel_err(got.cpu().numpy(), z["concat.y"]) <= 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("backend", ["simt", "auto"])
@pytest.mark.parametrize("shape,P", [((3, 256, 24, 40), 11), ((2, 256, 23, 37), 11), ((2, 128, 12, 20), 5), ((1, 64, 3, 5), 11)])
def test_correlate_concat_padded_layout(cuda_device, dtype, backend, shape, P):
    """The B200 concat layout [corr P*P | zeros | T2S_ref | T2S_next] (16-byte aligned blocks, the fast epilogue
    of the tcgen05 kernel) against the oracle's reference-order concat; pad channels must be exactly zero."""
    from stmask_b200.temporal_fusion import padded_corr_channels, unpad_concat
    ops = _ops()
    rng = np.random.default_rng(shape[2] * 7 + P)
    x1, x2 = q(rng.standard_normal(shape), dtype), q(rng.standard_normal(shape), dtype)
    fshape = (shape[0], 256, shape[2], shape[3])
    ta, tb = q(rng.standard_normal(fshape), dtype), q(rng.standard_normal(fshape), dtype)
    cp = padded_corr_channels(P)
    got = ops.correlation(dev(x1, dtype, cuda_device, True), dev(x2, dtype, cuda_device, True), P, 1, scale=1.0 / shape[1], relu=True,
                          feats=(dev(ta, dtype, cuda_device, True), dev(tb, dtype, cuda_device, True)), channels_last=True,
                          feat_channel_offset=cp, backend=backend)
    assert got.shape == (shape[0], cp + 512, shape[2], shape[3]) and got.stride(1) == 1
    assert (got[:, P * P:cp] == 0).all()
    want = np.maximum(np.concatenate([oracle.correlate(x1, x2, P, 1), ta, tb], 1), 0)    # relu(leaky(v)) == relu(v)
    assert rel_err(unpad_concat(got, P).float().cpu().numpy(), want) <= TOL[dtype]
    # the feature blocks are pure copies (+ReLU): bit exact
    assert torch.equal(got[:, cp:cp + 256].float().cpu(), torch.from_numpy(np.maximum(ta, 0)))
    assert torch.equal(got[:, cp + 256:].float().cpu(), torch.from_numpy(np.maximum(tb, 0)))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("backend", ["simt", "auto"])
@pytest.mark.parametrize("shape,P,d", [((2, 256, 24, 40), 11, 1), ((2, 256, 24, 40), 11, 2), ((1, 256, 48, 80), 11, 1),
                                       ((3, 256, 3, 5), 11, 1), ((2, 256, 6, 10), 11, 2), ((2, 256, 23, 40), 11, 1),
                                       ((2, 64, 12, 20), 5, 1), ((1, 40, 7, 9), 3, 3), ((1, 13, 6, 7), 7, 1),
                                       ((1, 512, 12, 20), 11, 1)])
def test_correlation_vs_oracle(cuda_device, dtype, backend, shape, P, d):
    from spatial_correlation_sampler import spatial_correlation_sample
    rng = np.random.default_rng(P * 100 + d + shape[2])
    x1, x2 = q(rng.standard_normal(shape), dtype), q(rng.standard_normal(shape), dtype)
    ops = _ops()
    got = ops.correlation(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), P, d, backend=backend,
                          out_dtype=torch.float32)
    want = oracle.correlation(x1, x2, P, d).reshape(shape[0], P * P, shape[2], shape[3])
    assert rel_err(got.cpu().numpy(), want) <= TOL[dtype]
    # 5-D drop-in API, NCHW-contiguous inputs
    got5 = spatial_correlation_sample(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), kernel_size=1, patch_size=P, stride=1,
                     padding=0, dilation_patch=d)
    assert got5.shape == (shape[0], P, P, shape[2], shape[3])
    assert rel_err(got5.float().cpu().numpy().reshape(want.shape), want) <= TOL[dtype]


def test_correlation_known_answers_full_size(cuda_device):
    ops = _ops()
    torch.manual_seed(4)
    # shift KAT at the sweep size: x2 = roll(x1, (+2, -1)) => peak at (ph, pw) = (7, 4) in the interior
    x1 = torch.randn(8, 256, 48, 80, device=cuda_device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x2 = torch.roll(x1, (2, -1), dims=(2, 3))
    out = ops.correlation(x1, x2, 11, 1, out_dtype=torch.float32)
    assert (out[:, :, 6:40, 6:70].argmax(1) == 7 * 11 + 4).all()
    # all ones => C where the displaced pixel is inside the map, 0 outside
    ones = torch.ones(2, 256, 24, 40, device=cuda_device, dtype=torch.bfloat16)
    out = ops.correlation(ones, ones, 11, 2, out_dtype=torch.float32).view(2, 11, 11, 24, 40)
    for ph in (0, 3, 5, 10):
        for pw in (0, 5, 9):
            dy, dx = (ph - 5) * 2, (pw - 5) * 2
            want = torch.zeros(24, 40, device=cuda_device)
            want[max(0, -dy):24 - max(0, dy), max(0, -dx):40 - max(0, dx)] = 256
            assert torch.equal(out[0, ph, pw], want)
    # symmetry: corr(x1, x2)[dy, dx](p) == corr(x2, x1)[-dy, -dx](p + d)
    a = torch.randn(4, 256, 24, 40, device=cuda_device)
    b = torch.randn(4, 256, 24, 40, device=cuda_device)
    ab = ops.correlation(a, b, 11, 1).view(4, 11, 11, 24, 40)
    ba = ops.correlation(b, a, 11, 1).view(4, 11, 11, 24, 40)
    assert rel_err(ab[:, 7, 4, 5:15, 8:30].cpu().numpy(), ba[:, 3, 6, 7:17, 7:29].cpu().numpy()) <= 1e-4


# ------------------------------------------------------------------------------------------
# error conventions: raise in Python before crossing the C ABI
# ------------------------------------------------------------------------------------------
def test_error_conventions(cuda_device):
    ops = _ops()
    x = torch.randn(1, 8, 5, 5, device=cuda_device)
    w = torch.randn(8, 8, 3, 3, device=cuda_device)
    with pytest.raises(ValueError):
        ops.deform_conv2d(x, torch.zeros(1, 18, 4, 4, device=cuda_device), w, padding=1)      # wrong offset size
    with pytest.raises(ValueError):
        ops.deform_conv2d(x[0], torch.zeros(1, 18, 5, 5, device=cuda_device), w, padding=1)   # not 4-D
    with pytest.raises(RuntimeError):
        ops.deform_conv2d(x.cpu(), torch.zeros(1, 18, 5, 5), w.cpu(), padding=1)              # CPU tensors
    with pytest.raises(TypeError):
        ops.deform_conv2d(x.half(), torch.zeros(1, 18, 5, 5, device=cuda_device), w.half(), padding=1)
    with pytest.raises(ValueError):
        ops.correlation(x, x, patch_size=4)
    with pytest.raises(ValueError):
        ops.correlation(x, x[:, :4], patch_size=3)
    from stmask_b200.compat.spatial_correlation_sampler import spatial_correlation_sample
    with pytest.raises(NotImplementedError):
        spatial_correlation_sample(x, x, kernel_size=3, patch_size=3)
    # batch sizes mmcv's im2col_step would reject (B % 32 != 0 for B > 32) work here
    xb = torch.randn(36, 8, 5, 5, device=cuda_device)
    assert ops.deform_conv2d(xb, torch.zeros(36, 18, 5, 5, device=cuda_device), w, padding=1).shape == (36, 8, 5, 5)
    # empty batch
    assert ops.deform_conv2d(x[:0], torch.zeros(0, 18, 5, 5, device=cuda_device), w, padding=1).shape == (0, 8, 5, 5)


# ------------------------------------------------------------------------------------------
# "identical downstream detections after fast NMS" (north star), on the reference's own head + NMS
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_detections_after_fast_nms_match_the_reference(cuda_device, dtype):
    """FCB(ada) 3x5 head (offsets -> deformable conv -> ReLU on this library; the plain class conv in torch fp32)
    on the inputs of tests/golden/detections.npz, then the reference's candidate filter + cross-class fast NMS
    (restated in conftest.detections_after_fast_nms and pinned to the reference's own output in test_oracle.py).
    fp32: the SAME detections (prior, class) with scores within 1e-4.
    bf16: sort / top-k / thresholds are discontinuous, so a 1e-2 feature error may flip near-ties (SURVEY.md
    hard part 5): detections whose reference score clears every decision by 0.02 must all be found with the same
    class and score within 2e-2, and at least 90 % of all detections must coincide."""
    import torch.nn.functional as F
    from conftest import detections_after_fast_nms
    from stmask_b200.feature_align import FeatureAlign
    z = load_golden("detections.npz")
    fa = FeatureAlign(64, 41, (3, 5), deformable_groups=1, use_pred_offset=True).to(cuda_device)
    with torch.no_grad():
        fa.conv_offset.weight.copy_(torch.from_numpy(z["w_offset"]))
        fa.conv_adaption.weight.copy_(torch.from_numpy(z["w_adaption"]))
    fa.conv_adaption.to(dtype)
    x = dev(z["x"], dtype, cuda_device, channels_last=True)
    with torch.no_grad():
        y = fa.calibrate_levels([x], [dev(z["shape"], torch.float32, cuda_device)])[0]
        logits = F.conv2d(y.float(), torch.from_numpy(z["w_conv"]).to(cuda_device), torch.from_numpy(z["b_conv"]).to(cuda_device),
                          padding=(1, 2)).cpu().numpy()
    assert rel_err(logits, z["logits"]) <= TOL[dtype]
    prior, cls, score = detections_after_fast_nms(logits[0], z["boxes"], z["centerness"])
    ref = {int(p): (int(c), float(s)) for p, c, s in zip(z["det_prior"], z["det_class"], z["det_score"])}
    got = {int(p): (int(c), float(s)) for p, c, s in zip(prior, cls, score)}
    if dtype == torch.float32:
        assert list(prior) == list(z["det_prior"]) and list(cls) == list(z["det_class"])
        assert np.abs(score - z["det_score"]).max() <= 1e-4
    else:
        common = set(ref) & set(got)
        assert len(common) >= 0.9 * max(len(ref), len(got)), (len(common), len(ref), len(got))
        assert all(ref[p][0] == got[p][0] and abs(ref[p][1] - got[p][1]) <= 2e-2 for p in common)


# ------------------------------------------------------------------------------------------
# RoIAlign (the step right after the temporal-fusion concat)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_roi_align_vs_golden_and_oracle(cuda_device, dtype):
    from mmcv.ops import roi_align                       # the drop-in the reference imports
    ops = _ops()
    z = load_golden("roi_align.npz")
    feat = q(z["tv.feat"], dtype)
    for name in ("a7", "a7_sr2", "l7", "a3x5_s05"):
        ph, pw, scale, sr, al = z[f"tv.{name}.cfg"]
        got = ops.roi_align(dev(feat, dtype, cuda_device), dev(z["tv.rois"], torch.float32, cuda_device), (int(ph), int(pw)),
                            float(scale), int(sr), bool(al))
        want = z[f"tv.{name}.y"] if dtype == torch.float32 else oracle.roi_align(feat, z["tv.rois"], (int(ph), int(pw)), float(scale), int(sr), bool(al))
        assert got.shape == want.shape and rel_err(got.float().cpu().numpy(), want) <= TOL[dtype], name
    # the reference's call site through the drop-in: roi_align(feature_maps, cat([box_ind, boxes]), 7), NCHW input
    rois = np.concatenate([np.zeros((9, 1), np.float32), z["ref.boxes"]], 1)
    f = q(z["ref.feat"], dtype)
    got = roi_align(dev(f, dtype, cuda_device), dev(rois, torch.float32, cuda_device), 7)
    want = z["ref.y"] if dtype == torch.float32 else oracle.roi_align(f, rois, 7)
    assert rel_err(got.float().cpu().numpy(), want) <= TOL[dtype]
    # on the padded 640-channel concat layout (channels-last, 16-byte vector path), many boxes, two images
    rng = np.random.default_rng(5)
    x = q(rng.standard_normal((2, 640, 24, 40)), dtype)
    b = rng.random((64, 4)).astype(np.float32)
    boxes = np.stack([b[:, 0] * 36, b[:, 1] * 20, b[:, 0] * 36 + 1 + b[:, 2] * 14, b[:, 1] * 20 + 1 + b[:, 3] * 10], 1)
    rois = np.concatenate([rng.integers(0, 2, (64, 1)).astype(np.float32), boxes], 1).astype(np.float32)
    got = ops.roi_align(dev(x, dtype, cuda_device, channels_last=True), dev(rois, torch.float32, cuda_device), 7)
    assert got.stride(1) == 1 and rel_err(got.float().cpu().numpy(), oracle.roi_align(x, rois, 7)) <= TOL[dtype]
    with pytest.raises(RuntimeError):
        ops.roi_align(torch.zeros(1, 8, 4, 4), torch.zeros(1, 5))
    with pytest.raises(ValueError):
        ops.roi_align(dev(x, dtype, cuda_device), torch.zeros(3, 4, device=cuda_device))


def test_correlation_pairs_reads_frames_and_halos_in_place(cuda_device):
    """Pair-indexed kernel (frames + received halos through index arrays) == the same kernel on gathered copies."""
    from stmask_b200 import sharding
    from stmask_b200.temporal_fusion import correlate_concat
    ops = _ops()
    torch.manual_seed(11)
    mk = lambda n, c: torch.randn(n, c, 24, 40, device=cuda_device).bfloat16().contiguous(memory_format=torch.channels_last)
    plan = sharding.make_plan(3, 6, 2, "frame")                 # rank 1 owns frames 3..5 of 3 clips: 3 halos, 9 pairs
    fpn, t2s = mk(9, 256), mk(9, 256)
    wire = torch.randn(3, 24, 40, 512, device=cuda_device).bfloat16()       # halos as they arrive: NHWC [fpn | t2s]
    halo_fpn, halo_t2s = wire.permute(0, 3, 1, 2)[:, :256], wire.permute(0, 3, 1, 2)[:, 256:]
    ref_idx, next_idx = sharding.pair_index_tensors(plan, 1, cuda_device)
    assert ref_idx.numel() == 9 and int(ref_idx.max()) >= 9
    got = ops.correlation_pairs(fpn, ref_idx, next_idx, 11, 1, scale=1 / 256, relu=True, feats=t2s, halo=halo_fpn,
                                feats_halo=halo_t2s, feat_channel_offset=128)
    fr, fn = sharding.temporal_pairs(plan, 1, fpn, halo_fpn)
    tr, tn = sharding.temporal_pairs(plan, 1, t2s, halo_t2s)
    want = correlate_concat(fr, fn, tr, tn, 11, 1, padded=True)
    assert got.shape == want.shape == (9, 640, 24, 40)
    assert torch.equal(got, want)
    # no features, no halo: plain indexed cost volume
    plan1 = sharding.make_plan(2, 4, 1, "clip")
    r1, n1 = sharding.pair_index_tensors(plan1, 0, cuda_device)
    x = mk(8, 128)
    got = ops.correlation_pairs(x, r1, n1, 11, 1)
    want = ops.correlation(x[r1.long()], x[n1.long()], 11, 1, channels_last=True)
    assert torch.equal(got, want)


# ------------------------------------------------------------------------------------------
# the hot path end to end from host memory == the device-resident step, bit for bit
# ------------------------------------------------------------------------------------------
def test_forward_streamed_equals_resident_step(cuda_device):
    from stmask_b200 import sharding
    from stmask_b200.hotpath import HotPath, HotPathConfig, StreamedIO
    cfg = HotPathConfig(backbone="r50", fcb="ada", height=96, width=160)
    hp = HotPath(cfg, cuda_device, seed=0)
    n = 10
    plan = sharding.make_plan(2, 5, 1, "clip")                 # two 5-frame clips: pairs must not cross the clip boundary
    d_in = hp.make_inputs(n, cuda_device, seed=3, on_device=True)
    want = dict(hp._frames_only({k: v for k, v in d_in.items() if not k.startswith("tf.")}))
    want.update(hp._tf_only({k: v for k, v in d_in.items() if k.startswith("tf.")}, plan, 0, None))
    whole = hp(d_in, plan, 0)                                  # the schedulable unit bench.py times
    assert want["tf.concat"].shape[0] == 8 and torch.equal(want["tf.concat"], whole["tf.concat"])
    # ... and the pair-indexed kernel equals the reference-order concat of gathered pairs
    from stmask_b200.temporal_fusion import correlate_concat, unpad_concat
    fr, fn = sharding.temporal_pairs(plan, 0, d_in["tf.fpn"], None)
    tr, tn = sharding.temporal_pairs(plan, 0, d_in["tf.t2s"], None)
    assert torch.equal(unpad_concat(whole["tf.concat"]), correlate_concat(fr, fn, tr, tn, channels_last=True, padded=True)[:, [*range(121), *range(128, 640)]])
    # end to end from pinned host slabs: 4 + 4 + 2 frames, one H2D and one D2H copy per chunk
    io = StreamedIO(hp, cuda_device, n, chunk_frames=4)
    host = io.host_buffers(plan.local_pairs(0))
    h_in, h_out, h_tin, h_tout = host
    for ci in range(io.n_chunks):
        a, b = ci * 4, min(n, ci * 4 + 4)
        for k, v in io.lin.views(h_in[ci], b - a).items():
            v.copy_(d_in[k][a:b])
    for k, v in io.tin.views(h_tin).items():
        v.copy_(d_in[k])
    torch.cuda.synchronize()
    for _ in range(2):                                         # second pass reuses the staging buffers
        for t in h_out:
            t.zero_()
        hp.forward_streamed(host, io, plan, 0)
    got = {"tf.concat": io.tout.views(h_tout)["tf.concat"]}
    for ci in range(io.n_chunks):
        a, b = ci * 4, min(n, ci * 4 + 4)
        for k, v in io.lout.views(h_out[ci], b - a).items():
            got.setdefault(k, []).append(v)
    for k, v in want.items():
        g = got[k] if k == "tf.concat" else torch.cat(got[k], 0)
        # every operator on the path (the DCN's offset/mask predictor included) is this library's own deterministic
        # kernel: chunked and whole-batch runs agree bit for bit
        assert torch.equal(g, v.cpu()), k


# ------------------------------------------------------------------------------------------
# host-side contracts: forward-only guard, packed-weight cache identity, pair-index validation
# ------------------------------------------------------------------------------------------
def test_forward_only_guard_raises_instead_of_cutting_the_graph(cuda_device):
    from dcn_v2 import DCN
    from mmcv.ops import DeformConv2d
    x = torch.randn(1, 16, 6, 6, device=cuda_device)
    off = torch.zeros(1, 18, 6, 6, device=cuda_device)
    m = DeformConv2d(16, 16, 3, padding=1).to(cuda_device)
    with pytest.raises(RuntimeError, match="forward-only"):
        m(x, off)                                        # weight requires grad and grad mode is on
    with pytest.raises(RuntimeError, match="forward-only"):
        DCN(16, 16, 3, 1, 1).to(cuda_device)(x)
    with torch.no_grad():
        assert m(x, off).shape == (1, 16, 6, 6)
    assert _ops().deform_conv2d(x.requires_grad_(False), off, m.weight.detach(), padding=1).shape == (1, 16, 6, 6)


def test_packed_weight_cache_cannot_alias_a_recycled_address(cuda_device):
    """A freed weight's address is routinely handed to the next same-shape tensor by the caching allocator; the
    functional entry points must not serve the old packed copy for it."""
    ops = _ops()
    torch.manual_seed(0)
    x = torch.randn(1, 32, 8, 8, device=cuda_device)
    off = torch.zeros(1, 18, 8, 8, device=cuda_device)
    seen = set()
    for i in range(6):
        w = torch.randn(32, 32, 3, 3, device=cuda_device)          # _version 0 every time
        seen.add(w.data_ptr())
        got = ops.deform_conv2d(x, off, w, padding=1)
        ref = torch.nn.functional.conv2d(x, w, padding=1)
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) <= 1e-4, i
        again = ops.deform_conv2d(x, off, w, padding=1)             # second call: served from the cache
        assert torch.equal(got, again)
        del w
    assert len(seen) < 6                                            # the allocator did recycle an address
    # `.data` writes are invisible to the version counter: documented, explicit invalidate()
    w = torch.randn(32, 32, 3, 3, device=cuda_device)
    cache = ops.PackedWeightCache()
    a = ops.deform_conv2d(x, off, w, padding=1, cache=cache)
    w.data.mul_(2.0)
    cache.invalidate()
    b = ops.deform_conv2d(x, off, w, padding=1, cache=cache)
    assert rel_err(b.cpu().numpy(), 2 * a.cpu().numpy()) <= 1e-6


def test_pair_indices_are_validated(cuda_device):
    ops = _ops()
    x = torch.randn(4, 64, 6, 10, device=cuda_device).bfloat16().contiguous(memory_format=torch.channels_last)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=cuda_device)
    with pytest.raises(ValueError, match="out of range"):
        ops.correlation_pairs(x, i32([0, 4]), i32([1, 2]), 11, 1)           # ref 4 >= 4 frames and no halo
    with pytest.raises(ValueError, match="out of range"):
        ops.correlation_pairs(x, i32([0, 1]), i32([1, 7]), 11, 1)
    with pytest.raises(ValueError, match="out of range"):
        ops.correlation_pairs(x, i32([-1, 1]), i32([1, 2]), 11, 1)
    # misaligned base with pair indexing: an error, not a silent fall-back to the un-indexed CUDA-core kernel
    from stmask_b200 import _lib
    flat = torch.randn(4 * 64 * 6 * 10 + 8, device=cuda_device).bfloat16()
    mis = flat[1:1 + 4 * 64 * 6 * 10].view(4, 6, 10, 64).permute(0, 3, 1, 2)     # 2-byte aligned NHWC view
    with pytest.raises(_lib.StmError, match="aligned"):
        ops.correlation_pairs(mis, i32([0, 1]), i32([1, 2]), 11, 1)


# ------------------------------------------------------------------------------------------
# TemporalNet: RoIAlign on the concat buffer -> conv1..3 (plain-conv mode of the tcgen05 main
# loop) -> AvgPool + fc + fc_coeff, against the reference's own TemporalNet.forward / bbox_feat_extractor
# ------------------------------------------------------------------------------------------
def _golden_temporal_net(device):
    from stmask_b200.temporal_net import TemporalNet
    z = load_golden("temporal_net.npz")
    torch.manual_seed(int(z["seed"]))
    net = TemporalNet(633)                                   # same seed, same construction order as the reference module
    cs = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in net.state_dict().values()])
    assert np.allclose(cs, z["checksums"], rtol=1e-9, atol=1e-9)       # summation order may differ, the weights may not
    return z, net.to(device)


def test_temporal_net_fp32_vs_reference(cuda_device):
    from stmask_b200.temporal_net import shift_candidates
    z, net = _golden_temporal_net(cuda_device)
    x_reg, x_coeff = net(dev(z["crops"], torch.float32, cuda_device))
    assert x_reg.shape == (5, 4) and x_coeff.shape == (5, 32)
    assert rel_err(x_reg.cpu().numpy(), z["x_reg"]) <= 1e-4 and rel_err(x_coeff.cpu().numpy(), z["x_coeff"]) <= 1e-4
    # with this library's RoIAlign in front (the reference's bbox_feat_extractor call site), reference channel layout
    x_reg, x_coeff = shift_candidates(net, dev(z["concat"], torch.float32, cuda_device, channels_last=True),
                                      dev(z["boxes"], torch.float32, cuda_device), torch.from_numpy(z["pair"]).to(cuda_device))
    assert rel_err(x_reg.cpu().numpy(), z["x_reg"]) <= 1e-4 and rel_err(x_coeff.cpu().numpy(), z["x_coeff"]) <= 1e-4


def test_temporal_net_bf16_padded_layout_on_tcgen05(cuda_device):
    """The hot path's form: the padded 640-channel channels-last concat (as the correlation kernel writes it) ->
    RoIAlign -> three tcgen05 plain convs -> pool + FC, bf16 storage; <= 1e-2 against the reference's fp32 outputs
    (inputs are bf16-exact; weights are rounded to bf16 by the module)."""
    from stmask_b200 import ops
    from stmask_b200.temporal_net import shift_candidates
    z, net = _golden_temporal_net(cuda_device)
    net = net.to(torch.bfloat16)
    c = torch.from_numpy(z["concat"])
    padded = torch.cat([c[:, :121], c.new_zeros(2, 7, 12, 20), c[:, 121:]], 1)
    pd = padded.to(cuda_device, torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert "tcgen05" in ops.deform_conv2d_variant([(5, 640, 7, 7)], ops.ConvSpec(640, 512, 3, 1, 1), torch.bfloat16, zero_offset=True)
    x_reg, x_coeff = shift_candidates(net, pd, dev(z["boxes"], torch.float32, cuda_device), torch.from_numpy(z["pair"]).to(cuda_device))
    assert x_reg.dtype == torch.float32
    assert rel_err(x_reg.cpu().numpy(), z["x_reg"]) <= 1e-2, rel_err(x_reg.cpu().numpy(), z["x_reg"])
    assert rel_err(x_coeff.cpu().numpy(), z["x_coeff"]) <= 1e-2, rel_err(x_coeff.cpu().numpy(), z["x_coeff"])


def test_temporal_net_many_boxes_tcgen05_vs_cuda_cores(cuda_device):
    """1500 boxes (73 500 GEMM rows: the paired 256-row plain-conv instantiation) — tcgen05 bf16 against the fp32
    CUDA-core path of the same module, which the golden test above pins to the reference."""
    from stmask_b200 import ops
    _, net = _golden_temporal_net(cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    x = torch.relu(torch.randn((1500, 7, 7, 640), generator=g, device=cuda_device)).bfloat16().permute(0, 3, 1, 2)
    x[:, 121:128] = 0
    v = ops.deform_conv2d_variant([tuple(x.shape)], ops.ConvSpec(640, 512, 3, 1, 1), torch.bfloat16, zero_offset=True)
    assert "plain=1" in v and "pair=1" in v and "rows=256" in v, v
    ref_reg, ref_coeff = net(torch.cat([x[:, :121], x[:, 128:]], 1).float())            # fp32, reference channel layout
    net16 = net.to(torch.bfloat16)
    got_reg, got_coeff = net16(x)
    # the fp32 run above used fp32 weights; the bf16 module rounds them — compare against fp32 math on the ROUNDED weights
    ref_reg, ref_coeff = net16.float()(torch.cat([x[:, :121], x[:, 128:]], 1).float())
    assert rel_err(got_reg.cpu().numpy(), ref_reg.cpu().numpy()) <= 1e-2
    assert rel_err(got_coeff.cpu().numpy(), ref_coeff.cpu().numpy()) <= 1e-2


# ------------------------------------------------------------------------------------------
# CUDA-graph capture and replay of the whole step (the ABI promises: no allocation, no sync, stream-ordered)
# ------------------------------------------------------------------------------------------
def test_hot_path_step_is_cuda_graph_capturable(cuda_device):
    from stmask_b200 import _lib, sharding
    from stmask_b200.hotpath import HotPath, HotPathConfig
    cfg = HotPathConfig(backbone="r50", fcb="ada", height=192, width=320)
    hp = HotPath(cfg, cuda_device, seed=1)
    plan = sharding.make_plan(2, 4, 1, "clip")
    inp = hp.make_inputs(8, cuda_device, seed=5, on_device=True)
    inp = {k: v.contiguous(memory_format=torch.channels_last) for k, v in inp.items()}
    side = torch.cuda.Stream(cuda_device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                            # warm-up: packs weights, fills the host-side caches
        for _ in range(2):
            eager = hp(inp, plan, 0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in eager.items()}
    graph = torch.cuda.CUDAGraph()
    n0 = _lib.launch_count()
    with torch.cuda.graph(graph):
        captured = hp(inp, plan, 0)
    n_kernels = _lib.launch_count() - n0
    assert n_kernels == 7 * 2 + 3 + 1                        # 7 DCN + 7 predictors, 3 grouped FCB launches (offsets fused), 1 TF
    graph.replay()
    torch.cuda.synchronize()
    for k, v in eager.items():
        assert torch.equal(captured[k], v), k
    # new data in the SAME input buffers, replay, compare with an eager run on that data
    fresh = hp.make_inputs(8, cuda_device, seed=6, on_device=True)
    for k in inp:
        inp[k].copy_(fresh[k])
    graph.replay()
    torch.cuda.synchronize()
    replayed = {k: v.clone() for k, v in captured.items()}
    eager2 = hp(inp, plan, 0)
    torch.cuda.synchronize()
    for k, v in eager2.items():
        assert torch.equal(replayed[k], v), k
    assert not torch.equal(replayed["tf.concat"], eager["tf.concat"])


def test_pack_modules_use_the_library_predictor(cuda_device):
    """mmcv.ops.DeformConv2dPack / ModulatedDeformConv2dPack with NON-zero offset predictors: their conv_offset runs on this library's
    own convolution (no cuDNN); result == the functional op fed with a torch fp32 conv2d of the same predictor."""
    from stmask_b200 import _lib as L
    from stmask_b200.compat.mmcv_ops import DeformConv2dPack, ModulatedDeformConv2dPack, deform_conv2d, modulated_deform_conv2d
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(4)
    x = torch.randn(2, 32, 12, 14, device=cuda_device)
    with torch.no_grad():
        for cls in (DeformConv2dPack, ModulatedDeformConv2dPack):
            m = cls(32, 48, 3, padding=1).to(cuda_device)
            torch.nn.init.normal_(m.conv_offset.weight, std=0.05)
            torch.nn.init.normal_(m.conv_offset.bias, std=0.3)
            n0 = L.launch_count()
            y = m(x)
            assert L.launch_count() - n0 >= 2                     # predictor + sampling kernel are both this library's
            om = torch.nn.functional.conv2d(x, m.conv_offset.weight, m.conv_offset.bias, padding=1)
            if cls is DeformConv2dPack:
                want = deform_conv2d(x, om, m.weight, 1, 1, 1, 1, 1)
            else:
                want = modulated_deform_conv2d(x, om[:, :18], torch.sigmoid(om[:, 18:]), m.weight, m.bias, 1, 1, 1, 1, 1)
            assert rel_err(y.cpu().numpy(), want.cpu().numpy()) <= 1e-4
