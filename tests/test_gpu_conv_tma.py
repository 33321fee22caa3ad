"""GPU parity of the TMA shifted-view convolution kernel (csrc/conv_tma.cu): the plain-conv (STM_DCN_ZERO_OFFSET) mode of
stm_deform_conv2d_fwd for shapes that tile well — the DCN offset/mask predictors (backbone.py:24-26), the prediction-head
convs (prediction_head_FC.py:71-127,150-183).  Reference: cuDNN fp32 convolution (TF32 off) on the bf16-rounded inputs;
and the gather main loop of the same library (STM_DCN_HINT_GATHER), which the oracle tests pin."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ref(x, w, b, stride, pad, relu):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = torch.nn.functional.conv2d(x.float(), w.float(), b, stride=stride, padding=pad)
    return torch.relu(y) if relu else y


CASES = [
    # (Cin, Cout, k, stride, pad, [(B, H, W), ...], out_f32, relu)
    (256, 32, (3, 3), 1, (1, 1), [(40, 24, 40)], True, False),          # backbone predictor, several tiles per CTA, resident weights
    (128, 32, (3, 3), 2, (1, 1), [(6, 96, 160)], True, False),          # stride-2 predictor (four parity tiles)
    (256, 32, (3, 3), 2, (1, 1), [(5, 48, 80)], True, False),
    (512, 32, (3, 3), 2, (1, 1), [(9, 24, 40)], True, False),           # weight ring (72 slices do not fit)
    (128, 32, (3, 3), 1, (1, 1), [(3, 48, 80)], True, False),
    (256, 256, (3, 3), 1, (1, 1), [(2, 48, 80), (2, 24, 40), (2, 12, 20), (2, 6, 10), (2, 3, 5)], False, True),   # head conv over P3..P7
    (256, 1024, (3, 3), 1, (1, 1), [(2, 24, 40)], False, True),         # four N tiles
    (256, 64, (3, 5), 1, (1, 2), [(3, 24, 40)], True, False),           # non-square kernels of the FCA head
    (256, 64, (5, 3), 1, (2, 1), [(3, 23, 37)], False, False),          # ragged map
    (64, 16, (1, 1), 1, (0, 0), [(2, 17, 29)], False, False),           # 1x1
    (192, 48, (1, 1), 1, (0, 0), [(21, 3, 5)], True, True),             # several images per tile
    (128, 32, (3, 3), 2, (1, 1), [(3, 25, 39)], True, False),           # stride 2 on odd sizes
    (256, 32, (3, 3), 1, (1, 1), [(3, 48, 80), (3, 11, 17), (5, 3, 5)], True, True),    # fused horizontal taps: ragged maps, several images per tile
    (64, 16, (3, 3), 1, (1, 1), [(2, 9, 130)], False, False),          # fused taps, a map wider than a tile
    (64, 32, (1, 1), 2, (0, 0), [(3, 24, 40)], False, False),           # 1x1 stride 2 (only the even / even parity tile is read)
    (64, 32, (3, 3), 2, (0, 0), [(3, 25, 41)], True, False),            # stride 2 without padding
    (64, 32, (5, 5), 2, (2, 2), [(2, 24, 40)], False, True),            # 5x5 stride 2: taps two lattice steps apart
    (64, 32, (3, 3), 1, (0, 0), [(3, 24, 40)], True, False),            # valid convolution (fused taps, no padding)
    (64, 32, (3, 3), 1, (2, 2), [(3, 12, 20)], True, False),            # padding larger than the kernel radius
    (128, 48, (5, 3), 2, (2, 1), [(2, 31, 47)], False, False),          # non-square stride 2 on odd sizes
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"C{c[0]}_N{c[1]}_k{c[2][0]}x{c[2][1]}_s{c[3]}_{len(c[5])}maps")
def test_tma_conv_vs_cudnn_fp32_and_gather_path(cuda_device, case):
    from stmask_b200 import ops
    from stmask_b200 import _lib as L
    cin, cout, k, s, pad, maps, out_f32, relu = case
    g = torch.Generator(device=cuda_device).manual_seed(cin + cout + k[0] * 7 + s)
    w = (torch.randn((cout, cin, *k), generator=g, device=cuda_device) / (cin * k[0] * k[1]) ** 0.5).bfloat16()
    b = torch.randn((cout,), generator=g, device=cuda_device)
    xs = [torch.randn((B, cin, H, W), generator=g, device=cuda_device).bfloat16().contiguous(memory_format=torch.channels_last)
          for (B, H, W) in maps]
    spec = ops.ConvSpec(cin, cout, k, s, pad)
    v = ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, zero_offset=True)
    assert "tma-conv" in v and "plain=1" in v, v
    vg = ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, zero_offset=True, hint=L.DCN_HINT_GATHER)
    assert "tma-conv" not in vg and "plain=1" in vg, vg
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    ys = ops.deform_conv2d_multi(xs, [None] * len(xs), [None] * len(xs), wp, b, spec, relu=relu, out_f32=out_f32)
    yg = ops.deform_conv2d_multi(xs, [None] * len(xs), [None] * len(xs), wp, b, spec, relu=relu, out_f32=out_f32,
                                 hint=L.DCN_HINT_GATHER)
    fused = s == 1 and k == (3, 3) and cout <= 32
    assert ("fused_taps=3" in v) == fused, v
    yn = ops.deform_conv2d_multi(xs, [None] * len(xs), [None] * len(xs), wp, b, spec, relu=relu, out_f32=out_f32, hint=L.DCN_HINT_NO_FUSE)
    assert "fused_taps=0" in ops.deform_conv2d_variant([tuple(x.shape) for x in xs], spec, torch.bfloat16, zero_offset=True, hint=L.DCN_HINT_NO_FUSE)
    torch.cuda.synchronize()
    for y, y3 in zip(ys, yn):
        assert rel_err(y.float().cpu().numpy(), y3.float().cpu().numpy()) <= (1e-4 if out_f32 else 1e-2)
    for x, y, y2 in zip(xs, ys, yg):
        want = _ref(x, w, b, s, pad, relu)
        assert y.shape == want.shape and y.dtype == (torch.float32 if out_f32 else torch.bfloat16)
        tol = 1e-4 if out_f32 else 1e-2
        assert rel_err(y.float().cpu().numpy(), want.cpu().numpy()) <= tol, rel_err(y.float().cpu().numpy(), want.cpu().numpy())
        assert rel_err(y.float().cpu().numpy(), y2.float().cpu().numpy()) <= tol


def test_tma_conv_writes_only_its_own_pixels(cuda_device):
    """The dead GEMM rows (halo columns, rows past the map) must never be stored: a strided output view keeps its guard bytes."""
    from stmask_b200 import ops
    g = torch.Generator(device=cuda_device).manual_seed(5)
    cin, cout = 128, 32
    x = torch.randn((3, cin, 13, 21), generator=g, device=cuda_device).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((cout, cin, 3, 3), generator=g, device=cuda_device) / 34.0).bfloat16()
    spec = ops.ConvSpec(cin, cout, 3, 1, 1)
    big = torch.full((3, 13, 21, 2 * cout), 7.0, device=cuda_device, dtype=torch.bfloat16)
    out = big[..., :cout].permute(0, 3, 1, 2)                      # NHWC view with a pixel stride of 2 * cout
    assert "tma-conv" in ops.deform_conv2d_variant([tuple(x.shape)], spec, torch.bfloat16, zero_offset=True)
    ops.deform_conv2d_multi([x], [None], [None], ops.pack_weight(w, spec, torch.bfloat16), None, spec, outs=[out])
    torch.cuda.synchronize()
    assert float((big[..., cout:] - 7.0).abs().max()) == 0.0
    want = _ref(x, w, None, 1, 1, False)
    assert rel_err(out.float().cpu().numpy(), want.cpu().numpy()) <= 1e-2


@pytest.mark.parametrize("case", [(256, 32, 1, [(5, 24, 40)]), (128, 32, 2, [(3, 48, 80)]), (256, 64, 1, [(2, 12, 20), (2, 6, 10)]),
                                  (640, 512, 1, [(3, 7, 7)])], ids=lambda c: f"C{c[0]}_N{c[1]}_s{c[2]}")
def test_plane_major_output_equals_channels_last(cuda_device, case):
    """STM_DCN_OUT_PLANAR: the same convolution written as contiguous [B, C, H, W] planes (fused-tap TMA kernel, plain TMA
    kernel, stride 2, and the gather main loop on 7x7 crops); fp32 and bf16 outputs."""
    from stmask_b200 import ops
    cin, cout, s, maps = case
    g = torch.Generator(device=cuda_device).manual_seed(cin + cout + s)
    w = (torch.randn((cout, cin, 3, 3), generator=g, device=cuda_device) / (cin * 9) ** 0.5).bfloat16()
    b = torch.randn((cout,), generator=g, device=cuda_device)
    xs = [torch.randn((B, cin, H, W), generator=g, device=cuda_device).bfloat16().contiguous(memory_format=torch.channels_last) for B, H, W in maps]
    spec = ops.ConvSpec(cin, cout, 3, s, 1)
    wp = ops.pack_weight(w, spec, torch.bfloat16)
    for f32 in (True, False):
        a = ops.deform_conv2d_multi(xs, [None] * len(xs), None, wp, b, spec, relu=True, out_f32=f32)
        p = ops.deform_conv2d_multi(xs, [None] * len(xs), None, wp, b, spec, relu=True, out_f32=f32, out_planar=True)
        torch.cuda.synchronize()
        for ya, yp in zip(a, p):
            assert yp.is_contiguous() and ya.is_contiguous(memory_format=torch.channels_last)
            assert torch.equal(ya, yp)
    # fp32 activations run on the CUDA-core kernel: channels-last result, same API
    y32 = ops.deform_conv2d_multi([xs[0].float()], [None], None, ops.pack_weight(w.float(), spec, torch.float32), b, spec, out_planar=True)[0]
    assert y32.is_contiguous(memory_format=torch.channels_last)
