"""Run the reference's OWN model (`/root/reference/STMask.py`, `STMask.forward_single` / `forward`, eval path with
Detect_TF -> Track_TF -> CandidateShift) on the CPU and record every hot-path operator call site.

TEST INFRASTRUCTURE ONLY (build container; /root/reference does not exist on the GPU box).  The three absent
third-party operators are replaced by the CPU stand-ins of `oracle/ref_harness.py`; everything else — backbone, FPN,
prediction heads, candidate generation, cross-class fast NMS, tracker, CandidateShift, TemporalNet — is the
reference's code, imported unmodified with the module shims of SURVEY.md §8(c):
  collections.Sequence alias (datasets/utils.py:2), torch.cuda.current_device patched (STMask.py:15; returns 'cpu'
  so TF_utils.py:104-109 works), stub modules dcn_v2, mmcv(.ops/.runner/.parallel), spatial_correlation_sampler,
  matplotlib(.pyplot/.patches/.collections), pycocotools(.mask/.coco/.cocoeval), cocoapi.PythonAPI.pycocotools.*,
  pyximport, utils.cython_nms.

BASELINE.json configs[0]: STMask R50-DCN-FPN FCA+TF (`STMask_plus_resnet50_config`), one synthetic 2-frame clip,
random weights, fp32.  The fixture uses a 96x160 frame (90x160 padded like 360x640 -> 384x640) so that the recorded
call-site tensors stay small; `eval_conf_thresh` / `nms_conf_thresh` are lowered so that the randomly initialised
heads produce candidates and the temporal-fusion path (correlate -> RoIAlign -> TemporalNet) runs on frame 2.
Parameters of the hot operators are overwritten from SEEDED generators (`seeded_params`), so the GPU test can rebuild
them without the reference: DCN weights U(+-1/sqrt(K)), biases N(0, 0.1^2), offset/mask predictor N(0, 0.05^2) /
N(0, 0.5^2) (non-zero: the reference zero-initialises them, which would hide sampling bugs), TemporalNet default init.
"""
from __future__ import annotations

import collections
import collections.abc
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import ref_harness as rh

CONFIG = "STMask_plus_resnet50_config"
FRAME_HW = (96, 160)
IMG_HW = (90, 160)
SEED = 20260202


def seeded_params(kind: str, index: int, shapes):
    """Deterministic parameter tensors for hot-path operator `index` of `kind`; the GPU test calls this too."""
    g = torch.Generator().manual_seed(SEED + {"dcn": 0, "temporal_net": 500}[kind] + index)
    out = []
    for name, shape in shapes:
        if name == "weight":
            k = shape[1] * shape[2] * shape[3]
            out.append((torch.rand(shape, generator=g) * 2 - 1) / k ** 0.5)
        elif name == "bias":
            out.append(torch.randn(shape, generator=g) * 0.1)
        elif name == "com_w":
            out.append(torch.randn(shape, generator=g) * 0.05)
        elif name == "com_b":
            out.append(torch.randn(shape, generator=g) * 0.5)
        else:
            raise KeyError(name)
    return out


def _install_stubs():
    collections.Sequence = collections.abc.Sequence
    torch.cuda.current_device = lambda: "cpu"

    def mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    mod("dcn_v2", DCN=rh.StubDCN, DCNv2=rh.StubDCNv2)
    ops = mod("mmcv.ops", DeformConv2d=rh.StubDeformConv2d, roi_align=rh._stub_roi_align)
    mm = mod("mmcv", ops=ops, is_str=lambda x: isinstance(x, str),
             is_list_of=lambda s, t: isinstance(s, list) and all(isinstance(i, t) for i in s))
    mm.__path__ = []
    mod("mmcv.runner", obj_from_dict=lambda *a, **k: None, get_dist_info=lambda: (0, 1))
    mod("mmcv.parallel", DataContainer=type("DataContainer", (), {}), collate=lambda b, **k: b)
    mod("spatial_correlation_sampler", spatial_correlation_sample=rh.stub_spatial_correlation_sample)
    mod("matplotlib").__path__ = []
    mod("matplotlib.pyplot")
    mod("matplotlib.patches", Polygon=None, Rectangle=None)
    mod("matplotlib.collections", PatchCollection=None)
    mod("pycocotools").__path__ = []
    mod("pycocotools.mask")
    mod("pycocotools.coco", COCO=None)
    mod("pycocotools.cocoeval", COCOeval=None)
    for n in ("cocoapi", "cocoapi.PythonAPI", "cocoapi.PythonAPI.pycocotools"):
        mod(n).__path__ = []
    mod("cocoapi.PythonAPI.pycocotools.ytvos", YTVOS=None)
    mod("cocoapi.PythonAPI.pycocotools.ytvoseval", YTVOSeval=None)
    mod("cocoapi.PythonAPI.pycocotools.mask")
    mod("pyximport", install=lambda *a, **k: None)
    try:
        import cv2  # noqa: F401
    except Exception:
        mod("cv2")
    if rh.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, rh.REFERENCE_ROOT)
    import utils  # noqa: F401  (the reference's package)
    mod("utils.cython_nms", nms=None)


def run_two_frame_clip():
    """-> dict of numpy arrays: the recorded call sites of frame 2 (and the clip-level inputs / results)."""
    _install_stubs()
    import STMask as S
    from datasets.config import cfg, set_cfg
    set_cfg(CONFIG)
    cfg.eval_conf_thresh = 0.012          # randomly initialised heads: class probabilities ~ 1/41, x centerness ~ 0.5
    cfg.nms_conf_thresh = 0.012
    cfg.max_num_detections = 12
    torch.manual_seed(SEED)
    torch.set_num_threads(1)
    net = S.STMask()
    net.eval()
    assert cfg.temporal_fusion_module and not cfg.use_dcn_class

    # ---- seeded parameters for the hot operators ----
    dcns = [m for m in net.modules() if isinstance(m, rh.StubDCN)]
    with torch.no_grad():
        for i, m in enumerate(dcns):
            w, b, cw, cb = seeded_params("dcn", i, [("weight", m.weight.shape), ("bias", m.bias.shape),
                                                    ("com_w", m.conv_offset_mask.weight.shape), ("com_b", m.conv_offset_mask.bias.shape)])
            m.weight.copy_(w); m.bias.copy_(b); m.conv_offset_mask.weight.copy_(cw); m.conv_offset_mask.bias.copy_(cb)
        torch.manual_seed(SEED + 500)
        fresh = type(net.TemporalNet)(net.TemporalNet.conv1.in_channels)          # default init under a recorded seed
        net.TemporalNet.load_state_dict(fresh.state_dict())

    rec = {}
    state = {"frame": 0}

    # ---- call-site recorders ----
    def wrap_dcn(i, m):
        orig = m.forward

        def fwd(x):
            y = orig(x)
            if state["frame"] == 1:
                rec[f"dcn{i}.x"], rec[f"dcn{i}.y"] = x.detach().clone(), y.detach().clone()
                rec[f"dcn{i}.stride"] = torch.tensor(m.stride[0])
            return y
        m.forward = fwd
    for i, m in enumerate(dcns):
        wrap_dcn(i, m)

    import layers.modules.track_to_segment_head as t2s
    import layers.functions.TF_utils as tfu
    orig_corr, orig_roi = t2s.spatial_correlation_sample, t2s.roi_align

    def rec_corr(x1, x2, **kw):
        out = orig_corr(x1, x2, **kw)
        rec["corr.x1"], rec["corr.x2"], rec["corr.out5d"] = x1.detach().clone(), x2.detach().clone(), out.detach().clone()
        return out

    def rec_roi(feat, rois, size, *a, **k):
        out = orig_roi(feat, rois, size, *a, **k)
        rec["roi.feat"], rec["roi.rois"], rec["roi.out"] = feat.detach().clone(), rois.detach().clone(), out.detach().clone()
        return out
    t2s.spatial_correlation_sample, t2s.roi_align = rec_corr, rec_roi
    orig_tn = net.TemporalNet.forward

    def rec_tn(x):
        a, b = orig_tn(x)
        rec["tn.x_reg"], rec["tn.x_coeff"] = a.detach().clone(), b.detach().clone()
        return a, b
    net.TemporalNet.forward = rec_tn
    orig_shift = tfu.CandidateShift

    def rec_shift(net_, ref_c, next_c, **kw):
        rec["shift.box_ref"] = ref_c["box"].detach().clone()
        rec["shift.t2s_ref"], rec["shift.t2s_next"] = ref_c["T2S_feat"].detach().clone(), next_c["T2S_feat"].detach().clone()
        out = orig_shift(net_, ref_c, next_c, **kw)
        rec["shift.box_ref_shift"] = out["box"].detach().clone()
        return out
    tfu.CandidateShift = rec_shift
    import layers.functions.track_TF as ttf
    ttf.CandidateShift = rec_shift

    H, W = FRAME_HW
    g = torch.Generator().manual_seed(SEED + 1)
    x1 = torch.randn(1, 3, H, W, generator=g)
    x1[:, :, IMG_HW[0]:] = 0                                   # impad_to_multiple zero padding (transforms.py:40-41)
    x2 = x1 + 0.1 * torch.randn(1, 3, H, W, generator=g)
    x2[:, :, IMG_HW[0]:] = 0

    def meta(first, fid):
        return {"img_shape": IMG_HW + (3,), "ori_shape": IMG_HW + (3,), "pad_shape": (H, W, 3), "is_first": first,
                "video_id": 0, "frame_id": fid, "scale_factor": 1.0, "flip": False}

    with torch.no_grad():
        state["frame"] = 0
        d1 = net(x1, [meta(True, 0)])[0]["detection"]
        state["frame"] = 1
        d2 = net(x2, [meta(False, 1)])[0]["detection"]
    t2s.spatial_correlation_sample, t2s.roi_align = orig_corr, orig_roi
    out = {k: v.numpy() for k, v in rec.items()}
    out["frames"] = torch.cat([x1, x2]).numpy()
    for name, d in (("det1", d1), ("det2", d2)):
        for k in ("box", "class", "score", "box_ids"):
            if k in d and torch.is_tensor(d[k]):
                out[f"{name}.{k}"] = d[k].detach().float().numpy()
    out["n_dcn"] = np.int64(len(dcns))
    return out
