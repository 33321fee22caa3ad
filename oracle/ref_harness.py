"""Import the reference's OWN Python call sites from /root/reference on top of CPU stand-ins
for the three third-party native operators.  TEST INFRASTRUCTURE ONLY; used by
``oracle/make_golden.py`` in the build container (``/root/reference`` does not exist on the
GPU box, so nothing here runs there).

Stand-ins:
  dcn_v2.DCN                      -> torchvision.ops.deform_conv2d (CPU) wrapped with the module
                                     semantics of CharlesShang/DCNv2 ``DCN.forward``
  mmcv.ops.DeformConv2d           -> torchvision.ops.deform_conv2d (CPU), no bias, (padH, padW)
  mmcv.ops.roi_align              -> torchvision.ops.roi_align(aligned=True)
  spatial_correlation_sampler     -> shifted-product definition in torch (fp64 accumulate)
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types
from contextlib import contextmanager

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("STM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "backbone.py"))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


# ----------------------------------------------------------------------------- stand-ins
class StubDCNv2(nn.Module):
    """DCNv2 base of CharlesShang/DCNv2 (dcn_v2.py): weight [Co,Ci,kh,kw], bias [Co]."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        n = in_channels * self.kernel_size[0] * self.kernel_size[1]
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.zero_()

    def forward(self, x, offset, mask):
        from torchvision.ops import deform_conv2d
        return deform_conv2d(x, offset, self.weight, self.bias, self.stride, self.padding, self.dilation, mask)


class StubDCN(StubDCNv2):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
        ch = self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(in_channels, ch, self.kernel_size, self.stride, self.padding, bias=True)
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, x):
        out = self.conv_offset_mask(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(mask)
        return StubDCNv2.forward(self, x, offset, mask)


class StubDeformConv2d(nn.Module):
    """mmcv.ops.DeformConv2d as the reference constructs it (Featurealign.py:27-31)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=False):
        super().__init__()
        assert not bias
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        n = in_channels * self.kernel_size[0] * self.kernel_size[1]
        self.weight.data.uniform_(-1.0 / math.sqrt(n), 1.0 / math.sqrt(n))

    def forward(self, x, offset):
        from torchvision.ops import deform_conv2d
        return deform_conv2d(x, offset, self.weight, None, self.stride, self.padding, self.dilation)


def stub_spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0,
                                    dilation=1, dilation_patch=1):
    assert kernel_size == 1 and stride == 1 and padding == 0 and dilation == 1
    B, C, H, W = input1.shape
    P, d = patch_size, dilation_patch
    r = P // 2
    pad = r * d
    x1 = input1.double()
    x2 = F.pad(input2.double(), (pad, pad, pad, pad))
    out = input1.new_zeros(B, P, P, H, W)
    for ph in range(P):
        for pw in range(P):
            dy, dx = (ph - r) * d + pad, (pw - r) * d + pad
            out[:, ph, pw] = (x1 * x2[:, :, dy:dy + H, dx:dx + W]).sum(1).to(input1.dtype)
    return out


def _stub_roi_align(feat, rois, out_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True):
    from torchvision.ops import roi_align
    return roi_align(feat, rois, _pair(out_size), spatial_scale, sampling_ratio, aligned)


@contextmanager
def _stubbed_modules(extra: dict):
    saved = {k: sys.modules.get(k) for k in extra}
    sys.modules.update(extra)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _load(relpath: str, name: str, stubs: dict):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    with _stubbed_modules(stubs):
        spec.loader.exec_module(mod)
    return mod


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def load_backbone():
    """reference backbone.py (ResNetBackbone, Bottleneck) over StubDCN."""
    return _load("backbone.py", "_stm_ref_backbone", {"dcn_v2": _mod("dcn_v2", DCN=StubDCN, DCNv2=StubDCNv2)})


def load_featurealign():
    """reference layers/modules/Featurealign.py (FeatureAlign) over StubDeformConv2d."""
    ops = _mod("mmcv.ops", DeformConv2d=StubDeformConv2d, roi_align=_stub_roi_align)
    mmcv = _mod("mmcv", ops=ops)
    return _load("layers/modules/Featurealign.py", "_stm_ref_featurealign", {"mmcv": mmcv, "mmcv.ops": ops})


def load_track_to_segment_head():
    """reference layers/modules/track_to_segment_head.py (correlate, TemporalNet, bbox_feat_extractor)."""
    ops = _mod("mmcv.ops", DeformConv2d=StubDeformConv2d, roi_align=_stub_roi_align)
    mmcv = _mod("mmcv", ops=ops)
    cfg = types.SimpleNamespace(use_sipmask=False, sipmask_head=4)

    def sanitize_coordinates_hw(box, h, w):
        # the reference's own layers/box_utils.py (the real module; box_utils needs only mmcv / utils / datasets stubs)
        return load_box_utils().sanitize_coordinates_hw(box, h, w)

    stubs = {
        "mmcv": mmcv, "mmcv.ops": ops,
        "spatial_correlation_sampler": _mod("spatial_correlation_sampler",
                                            spatial_correlation_sample=stub_spatial_correlation_sample),
        "datasets": _mod("datasets"), "datasets.config": _mod("datasets.config", cfg=cfg),
        "layers": _mod("layers"), "layers.box_utils": _mod("layers.box_utils",
                                                           sanitize_coordinates_hw=sanitize_coordinates_hw),
    }
    return _load("layers/modules/track_to_segment_head.py", "_stm_ref_t2s", stubs)


_BOX_UTILS = None


def load_box_utils():
    """reference layers/box_utils.py (jaccard, sanitize_coordinates_hw, ...)."""
    global _BOX_UTILS
    if _BOX_UTILS is None:
        from contextlib import contextmanager as _cm

        @_cm
        def _env(name):
            yield

        timer = _mod("utils.timer", env=_env)
        base = {"utils": _mod("utils", timer=timer), "utils.timer": timer,
                "datasets": _mod("datasets", cfg=types.SimpleNamespace(nms_as_miou=False)), "mmcv": _mod("mmcv")}
        _BOX_UTILS = _load("layers/box_utils.py", "_stm_ref_box_utils", base)
    return _BOX_UTILS


def load_detect(nms_as_miou: bool = False):
    """reference layers/functions/detection_TF.py (Detect_TF.detect / cc_fast_nms) with its own
    layers/box_utils.py (jaccard); masks are not involved (cfg.nms_as_miou = False as in the STMask configs)."""
    from contextlib import contextmanager as _cm

    @_cm
    def _env(name):
        yield

    cfg = types.SimpleNamespace(nms_as_miou=nms_as_miou)
    timer = _mod("utils.timer", env=_env)
    utils = _mod("utils", timer=timer)
    datasets = _mod("datasets", cfg=cfg)
    base = {"utils": utils, "utils.timer": timer, "datasets": datasets, "mmcv": _mod("mmcv")}
    box_utils = _load("layers/box_utils.py", "layers.box_utils", base)
    layers = _mod("layers", box_utils=box_utils)
    layers.__path__ = []
    functions = _mod("layers.functions")
    functions.__path__ = []
    mask_utils = _mod("layers.mask_utils", generate_mask=None)
    stubs = dict(base)
    stubs.update({"layers": layers, "layers.box_utils": box_utils, "layers.mask_utils": mask_utils, "layers.functions": functions})
    return _load("layers/functions/detection_TF.py", "layers.functions.detection_TF", stubs)
