"""BASELINE.json configs[0] — replay of the reference MODEL's hot-op call sites on the GPU (SURVEY.md 8(c) "Model:" KAT).

tests/golden/model_r50.npz holds what the reference's own `STMask.forward` (R50-DCN-FPN FCA+TF, STMask.py:205-329,
run on the CPU through oracle/ref_model.py over torchvision stand-ins) fed to and got from every operator on the hot
path for frame 2 of a synthetic 2-frame clip: the 7 backbone `DCN` modules (backbone.py:45), `correlate` /
`spatial_correlation_sample` (track_to_segment_head.py:53), the concat + ReLU (TF_utils.py:30-31), `roi_align`
(track_to_segment_head.py:86), `TemporalNet` (TF_utils.py:37) and the shifted boxes.  Each site is replayed through the
drop-in packages the reference would import (`dcn_v2`, `spatial_correlation_sampler`, `mmcv.ops` from shims/) with
the same seeded parameters: fp32 <= 1e-4 against the recorded outputs; bf16 <= 1e-2 against the oracle run on
bf16-rounded inputs.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden, rel_err
from oracle import ref_model

pytestmark = pytest.mark.gpu


def q(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dtype).float().numpy()


def dev(a, dtype, device, cl=False):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device=device, dtype=dtype)
    return t.contiguous(memory_format=torch.channels_last) if cl and t.dim() == 4 else t


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_backbone_dcn_call_sites(cuda_device, dtype):
    from dcn_v2 import DCN                                   # the reference's import (backbone.py:5), served by shims/
    z = load_golden("model_r50.npz")
    assert int(z["n_dcn"]) == 7
    for i in range(7):
        x, want = z[f"dcn{i}.x"], z[f"dcn{i}.y"]
        c, s = x.shape[1], int(z[f"dcn{i}.stride"])
        w, b, cw, cb = ref_model.seeded_params("dcn", i, [("weight", (c, c, 3, 3)), ("bias", (c,)), ("com_w", (27, c, 3, 3)), ("com_b", (27,))])
        m = DCN(c, c, kernel_size=3, stride=s, padding=1, dilation=1, deformable_groups=1).to(cuda_device)   # backbone.py:21-22
        with torch.no_grad():
            m.weight.copy_(w); m.bias.copy_(b); m.conv_offset_mask.weight.copy_(cw); m.conv_offset_mask.bias.copy_(cb)
            m = m.to(dtype)
            y = m(dev(x, dtype, cuda_device)).float().cpu().numpy()
        if dtype == torch.float32:
            assert rel_err(y, want) <= 1e-4, (i, rel_err(y, want))
        else:
            xr, wr, br, cwr, cbr = (q(t, dtype) for t in (x, w.numpy(), b.numpy(), cw.numpy(), cb.numpy()))
            ho, wo = want.shape[2:]
            om = oracle.deform_conv2d(xr, np.zeros((1, 18, ho, wo), np.float32), cwr, cbr, None, stride=s, padding=1)
            mask = (1.0 / (1.0 + np.exp(-om[:, 18:].astype(np.float64)))).astype(np.float32)
            ref = oracle.deform_conv2d(xr, om[:, :18], wr, br, mask, stride=s, padding=1)
            assert rel_err(y, ref) <= 1e-2, (i, rel_err(y, ref))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_temporal_fusion_call_sites(cuda_device, dtype):
    from mmcv.ops import roi_align                           # track_to_segment_head.py:6
    from spatial_correlation_sampler import spatial_correlation_sample      # track_to_segment_head.py:4
    from stmask_b200.temporal_fusion import correlate, correlate_concat
    from stmask_b200.temporal_net import TemporalNet, shift_candidates
    z = load_golden("model_r50.npz")
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    x1, x2 = q(z["corr.x1"], dtype), q(z["corr.x2"], dtype)
    ta, tb = q(z["shift.t2s_ref"], dtype), q(z["shift.t2s_next"], dtype)
    # correlate(): the reference's call with its keyword arguments
    out5 = spatial_correlation_sample(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), kernel_size=1, patch_size=11,
                                      stride=1, padding=0, dilation_patch=1)
    want5 = z["corr.out5d"] if dtype == torch.float32 else oracle.correlation(x1, x2, 11, 1)
    assert out5.shape == (1, 11, 11, 6, 10) and rel_err(out5.float().cpu().numpy(), want5) <= tol
    xc = correlate(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), 11)
    assert rel_err(xc.float().cpu().numpy(), oracle.correlate(x1, x2, 11, 1)) <= tol
    # relu(cat[x_corr, T2S_ref, T2S_next]) in one kernel == the tensor the reference handed to bbox_feat_extractor
    cat = correlate_concat(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), dev(ta, dtype, cuda_device), dev(tb, dtype, cuda_device))
    want_cat = z["roi.feat"] if dtype == torch.float32 else np.maximum(np.concatenate([oracle.correlate(x1, x2, 11, 1), ta, tb], 1), 0)
    assert cat.shape == (1, 633, 6, 10) and rel_err(cat.float().cpu().numpy(), want_cat) <= tol
    # roi_align(feature_maps, rois, 7) as bbox_feat_extractor calls it
    feat = q(z["roi.feat"], dtype)
    crops = roi_align(dev(feat, dtype, cuda_device), dev(z["roi.rois"][:8], torch.float32, cuda_device), 7)
    want_crops = z["roi.out"] if dtype == torch.float32 else oracle.roi_align(feat, z["roi.rois"][:8], 7)
    assert rel_err(crops.float().cpu().numpy(), want_crops) <= tol
    # the whole CandidateShift arithmetic on the device: concat (padded layout for bf16) -> RoIAlign -> TemporalNet -> decode
    torch.manual_seed(ref_model.SEED + 500)
    net = TemporalNet(633).to(cuda_device)
    boxes = dev(z["shift.box_ref"], torch.float32, cuda_device)
    if dtype == torch.bfloat16:
        net = net.to(dtype)
        cat = correlate_concat(dev(x1, dtype, cuda_device), dev(x2, dtype, cuda_device), dev(ta, dtype, cuda_device),
                               dev(tb, dtype, cuda_device), padded=True)
        assert cat.shape == (1, 640, 6, 10)
    x_reg, x_coeff = shift_candidates(net, cat, boxes, torch.zeros(boxes.shape[0], dtype=torch.long, device=cuda_device))
    assert rel_err(x_reg.cpu().numpy(), z["tn.x_reg"]) <= tol, rel_err(x_reg.cpu().numpy(), z["tn.x_reg"])
    assert rel_err(x_coeff.cpu().numpy(), z["tn.x_coeff"]) <= tol, rel_err(x_coeff.cpu().numpy(), z["tn.x_coeff"])
    # decode(loc, center_size(box_ref)) (box_utils.py:238-283, TF_utils.py:38): the shifted boxes the tracker consumes
    b = boxes.cpu()
    pri = torch.cat([(b[:, 2:] + b[:, :2]) / 2, b[:, 2:] - b[:, :2]], 1)
    loc = x_reg.cpu()
    dec = torch.cat([pri[:, :2] + loc[:, :2] * 0.1 * pri[:, 2:], pri[:, 2:] * torch.exp(loc[:, 2:] * 0.2)], 1)
    dec[:, :2] -= dec[:, 2:] / 2
    dec[:, 2:] += dec[:, :2]
    assert np.abs(dec.numpy() - z["shift.box_ref_shift"]).max() <= (1e-5 if dtype == torch.float32 else 2e-3)


# ------------------------------------------------------------------------------------------
# FCA / FCB prediction head over P3..P7 as grouped tcgen05 launches (SURVEY.md 8f rank 3, and rank 2's prior cache)
# ------------------------------------------------------------------------------------------
HEAD_KEYS = ("loc", "centerness", "conf", "mask_coeff", "track")


def _golden_head(device):
    from stmask_b200.prediction_head import PredictionHeadFC
    z = load_golden("prediction_head.npz")
    torch.manual_seed(int(z["seed"]))
    head = PredictionHeadFC(256, 41, 32, 128, fcb="ada")
    cs = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in head.state_dict().values()])
    assert np.allclose(cs, z["checksums"], rtol=1e-9, atol=1e-9)
    xs = [z[f"x{i}"] for i in range(5)]
    return z, head.to(device), xs


def test_prediction_head_fp32_vs_reference(cuda_device):
    z, head, xs = _golden_head(cuda_device)
    out = head.forward_levels([dev(x, torch.float32, cuda_device, cl=True) for x in xs])
    for k in HEAD_KEYS:
        assert out[k].shape == z[k].shape, k
        assert rel_err(out[k].cpu().numpy(), z[k]) <= 1e-4, (k, rel_err(out[k].cpu().numpy(), z[k]))
    assert np.array_equal(out["priors"].cpu().numpy(), z["priors"])
    assert rel_err(out["T2S_feat"][0].cpu().numpy(), z["T2S_feat0"]) <= 1e-4


def test_prediction_head_bf16_grouped_tcgen05_launches(cuda_device):
    """bf16: every conv of the head is ONE tcgen05 launch over the five levels (24 launches in all); <= 1e-2 against
    fp32 math on the same bf16-rounded weights (the CUDA-core path, pinned to the reference by the test above)."""
    from stmask_b200 import _lib
    z, head, xs = _golden_head(cuda_device)
    head16 = head.to(torch.bfloat16)
    xd = [dev(x, torch.bfloat16, cuda_device, cl=True) for x in xs]
    head16.forward_levels(xd)                                       # packs the weights
    n0 = _lib.launch_count()
    got = head16.forward_levels(xd)
    assert _lib.launch_count() - n0 == 1 + 1 + 4 + 3 * 5
    ref = head16.float().forward_levels([t.float() for t in xd])
    for k in HEAD_KEYS:
        assert rel_err(got[k].cpu().numpy(), ref[k].cpu().numpy()) <= 1e-2, (k, rel_err(got[k].cpu().numpy(), ref[k].cpu().numpy()))


# ------------------------------------------------------------------------------------------
# batched, sync-free candidate generation + cross-class fast NMS
# ------------------------------------------------------------------------------------------
def _to_center(b):
    return np.concatenate([(b[:, 2:] + b[:, :2]) / 2, b[:, 2:] - b[:, :2]], 1)


def test_detect_fast_nms_matches_the_reference_detections(cuda_device):
    """The reference's own post-NMS detections (tests/golden/detections.npz: its FeatureAlign head -> generate_candidate
    filter -> Detect_TF.cc_fast_nms) from the device kernel: same priors, classes and scores, in the same order."""
    from stmask_b200 import ops
    z = load_golden("detections.npz")
    logits = torch.from_numpy(z["logits"])[0]                                   # [41, h, w]
    c, h, w = logits.shape
    conf = torch.softmax(logits.reshape(c, h * w).t(), -1)[None].to(cuda_device)             # [1, P, 41]
    pri = torch.from_numpy(_to_center(z["boxes"])).to(cuda_device)               # decode(loc = 0, prior) == the recorded boxes
    loc = torch.zeros(1, h * w, 4, device=cuda_device)
    ctr = torch.from_numpy(z["centerness"]).reshape(1, -1).to(cuda_device)
    count, index, cls, score, box = ops.detect_fast_nms(conf, loc, ctr, pri, conf_thresh=0.05, nms_thresh=0.5, top_k=200)
    n = int(count[0])
    assert n == len(z["det_prior"])
    assert index[0, :n].cpu().tolist() == z["det_prior"].tolist()
    assert cls[0, :n].cpu().tolist() == z["det_class"].tolist()
    assert np.abs(score[0, :n].cpu().numpy() - z["det_score"]).max() <= 1e-6
    assert np.abs(box[0, :n].cpu().numpy() - z["boxes"][z["det_prior"]]).max() <= 1e-5


def test_detect_fast_nms_batched_full_size_vs_restatement(cuda_device):
    """4 frames x 15 345 priors x 41 classes (a 384x640 frame's P3..P7 with 3 anchors) in ONE launch against the
    restatement of the reference's filter + NMS (conftest.detections_after_fast_nms, pinned to the reference in test_oracle.py)."""
    from conftest import detections_after_fast_nms
    from stmask_b200 import ops
    from stmask_b200.prediction_head import make_priors
    g = torch.Generator().manual_seed(17)
    pri = torch.cat([make_priors(hh, ww) for hh, ww in ((48, 80), (24, 40), (12, 20), (6, 10), (3, 5))], 1)[0]
    P = pri.shape[0]
    assert P == 15345
    F = 4
    logits = torch.randn(F, P, 41, generator=g) * 2.0
    logits[:, :, 0] += 3.0                                                    # mostly background, a few hundred candidates
    conf = torch.softmax(logits, -1)
    loc = torch.randn(F, P, 4, generator=g) * 0.5
    ctr = torch.rand(F, P, generator=g)
    from stmask_b200 import _lib
    n0 = _lib.launch_count()
    count, index, cls, score, box = ops.detect_fast_nms(conf.to(cuda_device), loc.to(cuda_device), ctr.to(cuda_device), pri.to(cuda_device),
                                                         conf_thresh=0.05, nms_thresh=0.5, top_k=200)
    assert _lib.launch_count() - n0 == 1
    for f in range(F):
        # reference decode (box_utils.py:238-283)
        b = torch.cat([pri[:, :2] + loc[f, :, :2] * 0.1 * pri[:, 2:], pri[:, 2:] * torch.exp(loc[f, :, 2:] * 0.2)], 1)
        b[:, :2] -= b[:, 2:] / 2
        b[:, 2:] += b[:, :2]
        # the restatement takes [C, h, w] logits of one level; feed it log-probabilities laid out as a 1 x P "map"
        p_ref, c_ref, s_ref = detections_after_fast_nms(torch.log(conf[f]).t().reshape(41, 1, P), b, ctr[f])
        n = int(count[f])
        assert n == len(p_ref) and n > 20, (f, n, len(p_ref))
        assert index[f, :n].cpu().tolist() == p_ref.tolist()
        assert cls[f, :n].cpu().tolist() == c_ref.tolist()
        assert np.abs(score[f, :n].cpu().numpy() - s_ref).max() <= 1e-5
        assert np.abs(box[f, :n].cpu().numpy() - b[p_ref].numpy()).max() <= 1e-5


# ------------------------------------------------------------------------------------------
# mask assembly + mask IoU on the device (the rest of SURVEY.md 8f rank 4 short of the tracker's state machine)
# ------------------------------------------------------------------------------------------
def test_mask_assembly_and_mask_iou_vs_reference(cuda_device):
    from stmask_b200 import ops
    z = load_golden("mask_assembly.npz")
    proto = torch.from_numpy(z["proto"]).to(cuda_device)
    F = 3                                                        # the same frame three times with different valid counts
    na, nb = z["a.coeff"].shape[0], z["b.coeff"].shape[0]
    rep = lambda a: torch.from_numpy(a).to(cuda_device)[None].repeat(F, *([1] * a.ndim)).contiguous()
    cnt_a = torch.tensor([na, na - 2, 0], dtype=torch.int32, device=cuda_device)
    masks_a, bits_a = ops.mask_assembly(rep(z["proto"]), rep(z["a.coeff"]), rep(z["a.boxes"]), cnt_a)
    masks_b, bits_b = ops.mask_assembly(rep(z["proto"]), rep(z["b.coeff"]), rep(z["b.boxes"]))
    assert masks_a.shape == (F, na, 24, 40)
    assert np.abs(masks_a[0].cpu().numpy() - z["a.masks"]).max() <= 1e-5
    assert np.abs(masks_b[2].cpu().numpy() - z["b.masks"]).max() <= 1e-5
    assert np.abs(masks_a[1, :na - 2].cpu().numpy() - z["a.masks"][:na - 2]).max() <= 1e-5
    assert float(masks_a[1, na - 2:].abs().max()) == 0.0 and float(masks_a[2].abs().max()) == 0.0     # past count[f]: untouched
    iou = ops.mask_iou_bits(bits_a, bits_b, cnt_a, None)
    assert np.abs(iou[0].cpu().numpy() - z["iou"]).max() <= 1e-6
    assert np.abs(iou[1, :na - 2].cpu().numpy() - z["iou"][:na - 2]).max() <= 1e-6 and float(iou[2].abs().max()) == 0.0
    # the bit planes are the thresholded masks
    want_bits = (torch.from_numpy(z["a.masks"]) > 0.5).reshape(na, -1)
    got = bits_a[0].cpu().numpy().view(np.uint32)
    unpacked = ((got[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None]) & 1).reshape(na, -1)[:, :960].astype(bool)
    assert np.array_equal(unpacked, want_bits.numpy())
