"""CPU: the drop-in boundary — the C-ABI library loads and exports exactly what include/stmask_b200.h
declares, the Python binding covers every export, the product never touches the oracle and has no
CPU fallback, and the host-side mirrors keep the reference's names/shapes."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, golden_meta, load_golden

HEADER = os.path.join(ROOT, "include", "stmask_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from stmask_b200 import build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    declared = _declared()
    assert len(declared) >= 14
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared, (set(declared) ^ set(exported))
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name)
    lib.stm_version.restype = ctypes.c_int
    from stmask_b200 import _lib as _b
    assert lib.stm_version() == _b.ABI_VERSION
    lib.stm_last_error.restype = ctypes.c_char_p
    assert lib.stm_last_error() == b""


def test_python_binding_covers_the_header():
    from stmask_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    handle = _lib.lib()
    assert handle.stm_version() == _lib.ABI_VERSION
    # struct layouts: sizes follow from the header's field lists (no padding surprises)
    assert ctypes.sizeof(_lib.StmDcnConv) == 16 * 4
    assert ctypes.sizeof(_lib.StmDcnProblem) == 5 * 4 + 4 + 8 * (4 + 5 + 5 + 4)
    assert ctypes.sizeof(_lib.StmCorrDesc) == 12 * 4 + 8 * 10 + 2 * 4 + 8 * 6 + 2 * 4 + 2 * 8 + 4 * 4 + 2 * 8 + 6 * 8


def test_abi_validates_arguments_without_a_gpu():
    from stmask_b200 import _lib
    lib = _lib.lib()
    conv = _lib.StmDcnConv(256, 256, 3, 3, 1, 1, 1, 1, 1, 1, 1, 3, 1, 0, 0, 0)     # 256 % 3 != 0
    assert lib.stm_dcn_packed_weight_bytes(ctypes.byref(conv)) == 0
    assert b"deform_groups" in lib.stm_last_error()
    conv.deform_groups = 4
    assert lib.stm_dcn_packed_weight_bytes(ctypes.byref(conv)) == 256 * 256 * 9 * 2
    assert lib.stm_last_error() == b""
    prob = (_lib.StmDcnProblem * 1)()
    prob[0].batch, prob[0].in_h, prob[0].in_w, prob[0].out_h, prob[0].out_w = 1, 8, 8, 7, 8
    rc = lib.stm_deform_conv2d_backend(ctypes.byref(conv), prob, 1)
    assert rc == -1 and b"out size" in lib.stm_last_error()
    d = _lib.StmCorrDesc()
    d.batch, d.h, d.w, d.c, d.patch, d.dilation_patch = 1, 4, 4, 8, 4, 1
    assert lib.stm_correlation_fwd(ctypes.byref(d), None, None, None, None, None, None) == -1
    assert b"odd" in lib.stm_last_error()
    assert lib.stm_device_supported(0) in (0, 1)


def test_product_never_imports_the_oracle_and_has_no_cpu_fallback():
    pkg = os.path.join(ROOT, "stmask_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "stm_oracle" not in text, f
                assert "torchvision.ops import deform_conv2d" not in text, f
    from stmask_b200 import ops
    x = torch.zeros(1, 8, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.deform_conv2d(x, torch.zeros(1, 18, 4, 4), torch.zeros(8, 8, 3, 3), padding=1)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.correlation(x, x, 3)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.fcb_ali_offsets(torch.zeros(1, 4, 4, 4), (3, 3))


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from stmask_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.StmError, match="no CPU or eager fallback"):
        _lib.lib()


def test_drop_in_names_and_parameter_shapes():
    import dcn_v2
    import mmcv.ops
    import spatial_correlation_sampler as scs
    m = dcn_v2.DCN(128, 128, kernel_size=3, stride=2, padding=1, dilation=1, deformable_groups=1)
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "weight": (128, 128, 3, 3), "bias": (128,), "conv_offset_mask.weight": (27, 128, 3, 3),
        "conv_offset_mask.bias": (27,)}
    assert m.conv_offset_mask.stride == (2, 2) and float(m.conv_offset_mask.weight.detach().abs().sum()) == 0.0
    d = mmcv.ops.DeformConv2d(256, 256, kernel_size=(3, 5), padding=(1, 2), deform_groups=4)
    assert tuple(d.weight.shape) == (256, 256, 3, 5) and d.deform_groups == 4
    with pytest.raises(AssertionError):
        mmcv.ops.DeformConv2d(8, 8, 3, bias=True)
    assert callable(mmcv.ops.roi_align) and callable(mmcv.ops.modulated_deform_conv2d)
    p = mmcv.ops.ModulatedDeformConv2dPack(16, 16, 3, padding=1, deform_groups=2)
    assert tuple(p.conv_offset.weight.shape) == (54, 16, 3, 3)
    s = scs.SpatialCorrelationSampler(kernel_size=1, patch_size=11, stride=1, padding=0, dilation_patch=2)
    assert s.patch_size == 11 and callable(scs.spatial_correlation_sample)
    from stmask_b200.feature_align import FeatureAlign
    fa = FeatureAlign(256, 41, kernel_size=(5, 3), deformable_groups=1, use_pred_offset=True)
    assert {k: tuple(v.shape) for k, v in fa.state_dict().items()} == {
        "conv_offset.weight": (30, 4, 1, 1), "conv_adaption.weight": (256, 256, 5, 3),
        "conv.weight": (41, 256, 5, 3), "conv.bias": (41,)}
    assert "conv_offset.weight" not in FeatureAlign(256, 41, (3, 3), 1, use_pred_offset=False).state_dict()


def test_dcn_placement_rule_matches_reference():
    from stmask_b200 import backbone_dcn as b
    z = load_golden("backbone_dcn.npz")
    want = golden_meta(z, "placement")
    assert [list(t) for t in b.dcn_placement(*b.RESNET50_DCN)] == want["r50"]
    assert [list(t) for t in b.dcn_placement(*b.RESNET101_DCN)] == want["r101"]
    assert b.dcn_placement([3, 4, 6, 3]) == [] == want["r50_nodcn"]
    assert len(want["r50"]) == 7 and len(want["r101"]) == 11
    shapes = b.dcn_layer_shapes(*b.RESNET101_DCN, 384, 640)
    assert [(s.channels, s.in_h, s.in_w, s.out_h, s.out_w) for s in shapes[:3]] == [
        (128, 96, 160, 48, 80), (128, 48, 80, 48, 80), (256, 48, 80, 24, 40)]
    assert all(s.flops_per_frame == 1132462080 for s in shapes)        # 1.1325 GF per layer
    assert (shapes[-1].channels, shapes[-1].out_h, shapes[-1].out_w) == (512, 12, 20)


def test_fake_impls_give_shapes_without_a_gpu():
    from stmask_b200 import ops  # noqa: F401  (registers the ops)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(2, 256, 24, 40, device="cuda")
        off = torch.empty(2, 30, 24, 40, device="cuda")
        w = torch.empty(256, 256, 3, 5, device="cuda")
        y = torch.ops.stmask_b200.deform_conv2d(x, off, None, w, None, [1, 1], [1, 2], [1, 1], 1, 1, True, False)
        assert tuple(y.shape) == (2, 256, 24, 40)
        c = torch.ops.stmask_b200.correlation(x, x, 11, 1, 1.0 / 256, 0.1, False)
        assert tuple(c.shape) == (2, 121, 24, 40)


def test_mmcv_shim_falls_through_to_a_real_mmcv(tmp_path):
    """shims/ ahead on PYTHONPATH must not shadow the rest of mmcv: a 'real' mmcv further down the path keeps
    serving imread / is_str / parallel / runner, only the hot-path names of mmcv.ops are replaced."""
    real = tmp_path / "site" / "mmcv"
    (real / "parallel").mkdir(parents=True)
    (real / "ops").mkdir()
    (real / "__init__.py").write_text("from .parallel import DataContainer\n__version__ = '1.1.2'\n"
                                      "def imread(p):\n    return ('imread', p)\ndef is_str(x):\n    return isinstance(x, str)\n")
    (real / "parallel" / "__init__.py").write_text("class DataContainer:\n    pass\ndef collate(b):\n    return b\n")
    (real / "ops" / "__init__.py").write_text("raise ImportError('mmcv._ext is not built')\n")      # mmcv-full without its extension
    code = ("import mmcv, mmcv.ops\n"
            "from mmcv.ops import DeformConv2d, roi_align\n"
            "from mmcv.parallel import DataContainer, collate\n"
            "assert mmcv.imread('x') == ('imread', 'x') and mmcv.is_str('a') and mmcv.__version__ == '1.1.2'\n"
            "assert DeformConv2d.__module__ == 'stmask_b200.compat.mmcv_ops' and mmcv.ops.__stmask_b200__\n"
            "import dcn_v2, spatial_correlation_sampler\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "shims"), ROOT, str(tmp_path / "site")]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
    # without any real mmcv: a namespace with mmcv.ops alone; and the PYTHONPATH-free form
    code2 = ("import stmask_b200\nstmask_b200.install_shims()\n"
             "from mmcv.ops import roi_align, ModulatedDeformConv2d\nfrom dcn_v2 import DCN\n"
             "from spatial_correlation_sampler import spatial_correlation_sample, SpatialCorrelationSampler\nprint('ok')\n")
    r = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT), cwd=str(tmp_path))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_launch_plan_is_a_pure_function_of_the_call():
    """stm_deform_conv2d_variant: which tcgen05 instantiation a call runs (no environment knobs; hints travel with the
    call).  Without a GPU the SM count defaults to 148 (B200)."""
    import torch
    from stmask_b200 import _lib, ops
    fpn = [(72, 256, h, w) for h, w in ((48, 80), (24, 40), (12, 20), (6, 10), (3, 5))]
    fcb = ops.ConvSpec(256, 256, (3, 5), 1, (1, 2))
    v = ops.deform_conv2d_variant(fpn, fcb, torch.bfloat16)
    assert "tcgen05 rows=128 n=256 pair=1 plain=0 fcb=0 producer_warps=8" in v and "ctas_per_sm=2" in v, v
    assert "rows=256 n=256 pair=0" in ops.deform_conv2d_variant(fpn, fcb, torch.bfloat16, hint=_lib.DCN_HINT_NO_PAIR)
    assert "rows=256 n=256 pair=1" in ops.deform_conv2d_variant(fpn, fcb, torch.bfloat16, hint=_lib.DCN_HINT_ROWS256)
    assert "rows=128 n=256 pair=1" in ops.deform_conv2d_variant(fpn, fcb, torch.bfloat16, hint=_lib.DCN_HINT_ROWS128)
    assert ops.deform_conv2d_variant(fpn, fcb, torch.float32) == "simt"
    small = ops.deform_conv2d_variant([(2, 256, 12, 20)], fcb, torch.bfloat16)
    assert "rows=128" in small and "pair=0" in small, small          # 480 rows: 4 CTAs, nothing to pair
    c128 = ops.deform_conv2d_variant([(72, 128, 48, 80)], ops.ConvSpec(128, 128, 3, 1, 1), torch.bfloat16)
    assert "rows=128 n=128" in c128 and "ctas_per_sm=2" in c128, c128
    pred = ops.deform_conv2d_variant([(72, 256, 24, 40)], ops.ConvSpec(256, 32, 3, 1, 1), torch.bfloat16, zero_offset=True)
    assert "plain=1" in pred and "n=32" in pred and "tma-conv stride=1" in pred and "resident=1" in pred, pred
    pred2 = ops.deform_conv2d_variant([(72, 512, 24, 40)], ops.ConvSpec(512, 32, 3, 2, 1), torch.bfloat16, zero_offset=True)
    assert "tma-conv stride=2" in pred2 and "resident=0" in pred2, pred2          # 72 weight slices: a ring, not resident
    gath = ops.deform_conv2d_variant([(72, 256, 24, 40)], ops.ConvSpec(256, 32, 3, 1, 1), torch.bfloat16, zero_offset=True,
                                     hint=_lib.DCN_HINT_GATHER)
    assert "tma-conv" not in gath and "plain=1" in gath, gath
    crops = ops.deform_conv2d_variant([(1500, 640, 7, 7)], ops.ConvSpec(640, 512, 3, 1, 1), torch.bfloat16, zero_offset=True)
    assert "tma-conv" not in crops, crops                            # 7x7 crops leave 62 % of a shifted-view tile dead: gather loop
    for k in ("STM_DCN_MTILES", "STM_DCN_PW", "STM_CORR_TW"):          # no hidden global state behind the ABI
        os.environ[k] = "1"
    try:
        assert ops.deform_conv2d_variant(fpn, fcb, torch.bfloat16) == v
    finally:
        for k in ("STM_DCN_MTILES", "STM_DCN_PW", "STM_CORR_TW"):
            os.environ.pop(k)
    src = "".join(open(os.path.join(ROOT, "stmask_b200", "csrc", f)).read() for f in os.listdir(os.path.join(ROOT, "stmask_b200", "csrc")))
    import re as _re
    stripped = _re.sub(r"#ifdef STM_DCN_EXPERIMENTS.*?#endif", "", src, flags=_re.S)
    assert "getenv" not in stripped


def test_tma_conv_plan_properties():
    """conv_tma.cu's launch plan over a sweep of shapes, queried without a GPU: the tile fits one 128-row accumulator (with the two
    extra rows the fused horizontal taps read), at least 55 % of its rows are live, and which kernel runs (TMA shifted views or
    the gather loop) never depends on how a caller chunks its batch — the two accumulate in different orders."""
    import re
    import torch
    from stmask_b200 import ops
    rx = re.compile(r"tma-conv stride=(\d) n=(\d+) fused_taps=(\d) tile=(\d+)x(\d+)x(\d+) live=([0-9.]+)")
    seen_tma = seen_gather = 0
    for cin, cout in ((64, 16), (128, 32), (256, 256)):
        for (kh, kw) in ((1, 1), (3, 3), (3, 5), (5, 3)):
            for s in (1, 2):
                for (h, w) in ((3, 5), (6, 10), (7, 7), (12, 20), (24, 40), (23, 37), (48, 80), (96, 160), (9, 130)):
                    spec = ops.ConvSpec(cin, cout, (kh, kw), s, (kh // 2, kw // 2))
                    ho, wo = spec.out_hw(h, w)
                    if ho <= 0 or wo <= 0:
                        continue
                    kinds = set()
                    for b in (1, 2, 3, 8, 40):
                        v = ops.deform_conv2d_variant([(b, cin, h, w)], spec, torch.bfloat16, zero_offset=True)
                        m = rx.search(v)
                        kinds.add(m is not None)
                        if m is None:
                            assert "plain=1" in v and "tma-conv" not in v, v
                            continue
                        stride, n, fused, bb, th, tw, live = int(m[1]), int(m[2]), int(m[3]), int(m[4]), int(m[5]), int(m[6]), float(m[7])
                        assert stride == s and n == min(cout, 256) and fused == (3 if (s == 1 and (kh, kw) == (3, 3) and cout <= 32) else 0), v
                        ex = (kw - 1) if s == 1 else (kw - 1 - kw // 2 - ((kw - 1 - kw // 2) & 1)) // 2 + (kw // 2 + 1) // 2
                        ey = (kh - 1) if s == 1 else (kh - 1 - kh // 2 - ((kh - 1 - kh // 2) & 1)) // 2 + (kh // 2 + 1) // 2
                        bw, bh = tw + ex, th + ey
                        rows = (bb - 1) * bh * bw + (th - 1) * bw + tw
                        assert rows <= 128 - (fused - 1 if fused else 0), (v, rows)
                        assert th <= ho and tw <= wo and bb <= b and 0.0 < live <= 1.0, v
                        if b == 40:
                            assert live >= 0.45, v          # the 55 % rule is applied to the batch-independent tile efficiency
                    assert len(kinds) == 1, (cin, cout, kh, kw, s, h, w)          # one kernel family whatever the batch size
                    seen_tma += True in kinds
                    seen_gather += False in kinds
    assert seen_tma > 50 and seen_gather > 5
