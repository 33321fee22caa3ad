"""Generate tests/golden/*.npz — run ONCE in the build container (needs /root/reference and
torchvision CPU); the fixtures are committed and are what the GPU box sees.

    python -m oracle.make_golden

Sources of truth:
  dcn_torchvision.npz   torchvision.ops.deform_conv2d (CPU, fp32): seeded random cases over kernel
                        shapes 3x3/3x5/5x3, stride 1/2, dilation, groups, deform_groups, mask, bias
  dcn_border.npz        border-rule known answers measured on torchvision
  feature_align.npz     the reference's own FeatureAlign.forward (layers/modules/Featurealign.py:42-74),
                        FCB(ada) and FCB(ali), imported from /root/reference over the torchvision stand-in
  correlate.npz         the reference's own correlate() (track_to_segment_head.py:40-62) over the
                        shifted-product stand-in, P=11, d=1/2
  detections.npz        FCB(ada) 3x5 head -> class confidences -> the reference's OWN candidate filter and
                        cross-class fast NMS (TF_utils.py:68-74, detection_TF.py:56-134): the inputs, the
                        reference's logits and the detections it keeps ("identical detections after fast NMS")
  roi_align.npz         torchvision.ops.roi_align (CPU) cases (aligned / legacy, adaptive / fixed sampling grid,
                        boxes over the border) and the reference's own bbox_feat_extractor call site
                        (track_to_segment_head.py:65-88) with its own sanitize_coordinates_hw
  model_r50.npz         BASELINE.json configs[0]: the reference's own STMask (R50-DCN-FPN FCA+TF, STMask.py:205-329) on a
                        synthetic 2-frame clip via oracle/ref_model.py: inputs / outputs of all 7 backbone DCN call sites,
                        the correlate / concat / RoIAlign / TemporalNet call sites of CandidateShift and the shifted boxes
  prediction_head.npz   the reference's own PredictionModule_FC.forward (R101 FCA+FCB(ada) head, shared over five levels):
                        loc / centerness / conf / mask_coeff / track / priors, weights rebuilt from the recorded seed
  tracker.npz           the reference's own Track_TF.track on a synthetic 7-frame clip (track_TF.py:52-181), recorded shifts
  mask_assembly.npz     the reference's own generate_mask / crop / mask_iou (mask_utils.py:111-128, box_utils.py:341-364,435-447)
  temporal_net.npz      the reference's own TemporalNet.forward + bbox_feat_extractor (track_to_segment_head.py:10-37,
                        65-88) on a 633-channel concat; the 40 MB of weights are rebuilt from the recorded seed
                        (default nn init in construction order) and pinned by per-parameter checksums
  backbone_dcn.npz      the reference's ResNetBackbone DCN placement (backbone.py:105-138) for the R50/R101
                        configs, and a Bottleneck DCN-branch forward (backbone.py:20-26,45)
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

DCN_CASES = [
    # name,            B, Cin, Cout, H,  W,  kh, kw, s, p,      d, groups, dg, mask,  bias
    ("k3x3",           2, 16,  24,   9,  11, 3,  3,  1, (1, 1), 1, 1,      1,  False, False),
    ("k3x5",           2, 16,  16,   8,  12, 3,  5,  1, (1, 2), 1, 1,      1,  False, False),
    ("k5x3",           2, 16,  16,   8,  12, 5,  3,  1, (2, 1), 1, 1,      1,  False, False),
    ("k5x3_on_3x5",    1, 16,  16,   3,  5,  5,  3,  1, (2, 1), 1, 1,      1,  False, False),
    ("k3x3_dg4",       2, 32,  16,   7,  9,  3,  3,  1, (1, 1), 1, 1,      4,  False, False),
    ("k3x5_dg4",       1, 32,  16,   6,  10, 3,  5,  1, (1, 2), 1, 1,      4,  False, False),
    ("v2_s1",          2, 16,  16,   9,  11, 3,  3,  1, (1, 1), 1, 1,      1,  True,  True),
    ("v2_s2",          2, 16,  16,   12, 16, 3,  3,  2, (1, 1), 1, 1,      1,  True,  True),
    ("v2_s2_odd",      1, 16,  24,   11, 15, 3,  3,  2, (1, 1), 1, 1,      1,  True,  True),
    ("v2_dil2",        1, 16,  16,   10, 10, 3,  3,  1, (2, 2), 2, 1,      1,  True,  True),
    ("v2_dg2_g2",      1, 16,  16,   8,  8,  3,  3,  1, (1, 1), 1, 2,      2,  True,  True),
    ("v1_c13_o7",      1, 13,  7,    6,  7,  3,  3,  1, (1, 1), 1, 1,      1,  False, False),
    ("v1_pad0",        1, 16,  16,   8,  9,  3,  3,  1, (0, 0), 1, 1,      1,  False, False),
]


def _dcn_torchvision():
    from torchvision.ops import deform_conv2d
    out = {}
    meta = {}
    for idx, (name, B, Cin, Cout, H, W, kh, kw, s, p, d, groups, dg, use_mask, use_bias) in enumerate(DCN_CASES):
        g = torch.Generator().manual_seed(1000 + idx)
        Ho = (H + 2 * p[0] - d * (kh - 1) - 1) // s + 1
        Wo = (W + 2 * p[1] - d * (kw - 1) - 1) // s + 1
        x = torch.randn(B, Cin, H, W, generator=g)
        offset = torch.randn(B, dg * 2 * kh * kw, Ho, Wo, generator=g) * 2.0      # N(0, 2^2) px
        weight = torch.randn(Cout, Cin // groups, kh, kw, generator=g) / (Cin // groups * kh * kw) ** 0.5
        mask = torch.rand(B, dg * kh * kw, Ho, Wo, generator=g) if use_mask else None
        bias = torch.randn(Cout, generator=g) if use_bias else None
        y = deform_conv2d(x, offset, weight, bias, (s, s), p, (d, d), mask)
        out[f"{name}.x"], out[f"{name}.offset"], out[f"{name}.weight"], out[f"{name}.y"] = (
            x.numpy(), offset.numpy(), weight.numpy(), y.numpy())
        if use_mask:
            out[f"{name}.mask"] = mask.numpy()
        if use_bias:
            out[f"{name}.bias"] = bias.numpy()
        meta[name] = dict(stride=s, padding=list(p), dilation=d, groups=groups, deform_groups=dg)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "dcn_torchvision.npz"), **out)


def _dcn_border():
    """1x1 kernel on a ramp image, one sample position per case (SURVEY.md Appendix A KATs)."""
    from torchvision.ops import deform_conv2d
    H, W = 5, 6
    x = torch.arange(1, H * W + 1, dtype=torch.float32).view(1, 1, H, W)
    w = torch.ones(1, 1, 1, 1)
    positions = [(-0.5, 0.0), (-0.99, 0.0), (-1.0, 0.0), (H - 0.5, 0.0), (float(H), 0.0), (H - 1.0, 0.0),
                 (0.0, -0.5), (0.0, -1.0), (0.0, W - 0.5), (0.0, float(W)), (1.5, 2.5), (-0.5, -0.5),
                 (H - 0.5, W - 0.5), (2.25, 3.75), (-3.0, 2.0), (2.0, 40.0)]
    vals = []
    for (h, wv) in positions:
        off = torch.zeros(1, 2, H, W)
        # output pixel (0,0) samples at (h, w): offset = target - base position (0,0)
        off[0, 0, 0, 0], off[0, 1, 0, 0] = h, wv
        y = deform_conv2d(x, off, w)
        vals.append(float(y[0, 0, 0, 0]))
    np.savez_compressed(os.path.join(OUT, "dcn_border.npz"), x=x.numpy(),
                        positions=np.asarray(positions, np.float32), values=np.asarray(vals, np.float32))


def _feature_align():
    fa_mod = rh.load_featurealign()
    out = {}
    C = 32
    cases = [("k3x3", (3, 3), 6, 10), ("k3x5", (3, 5), 6, 10), ("k5x3", (5, 3), 6, 10), ("k5x3_p7", (5, 3), 3, 5)]
    for mode in ("ada", "ali"):
        for idx, (name, ks, H, W) in enumerate(cases):
            torch.manual_seed(2000 + idx + (100 if mode == "ali" else 0))
            m = fa_mod.FeatureAlign(C, 41, kernel_size=ks, deformable_groups=1, use_pred_offset=(mode == "ada"))
            if mode == "ada":
                torch.nn.init.normal_(m.conv_offset.weight, std=0.5)    # non-zero
            torch.nn.init.normal_(m.conv_adaption.weight, std=(C * ks[0] * ks[1]) ** -0.5)
            x = torch.randn(2, C, H, W)
            shape = torch.randn(2, 4, H, W)
            cap = {}
            hk = m.conv_adaption.register_forward_hook(
                lambda mod, inp, o, cap=cap: cap.update(offset=inp[1].detach().clone(), dcn=o.detach().clone()))
            with torch.no_grad():
                y = m(x.clone(), shape)
            hk.remove()
            key = f"{mode}.{name}"
            out[f"{key}.x"], out[f"{key}.shape"] = x.numpy(), shape.numpy()
            out[f"{key}.w_adaption"] = m.conv_adaption.weight.detach().numpy()
            out[f"{key}.w_conv"], out[f"{key}.b_conv"] = m.conv.weight.detach().numpy(), m.conv.bias.detach().numpy()
            if mode == "ada":
                out[f"{key}.w_offset"] = m.conv_offset.weight.detach().numpy()
            out[f"{key}.offset"] = cap["offset"].numpy()
            # conv_adaption's output is ReLU-ed in place by the reference (nn.ReLU(inplace=True)),
            # the hook clones BEFORE that happens
            out[f"{key}.dcn_relu"] = torch.relu(cap["dcn"]).numpy()
            out[f"{key}.y"] = y.numpy()
    np.savez_compressed(os.path.join(OUT, "feature_align.npz"), **out)


def _correlate():
    t2s = rh.load_track_to_segment_head()
    out = {}
    for idx, (name, C, H, W, P, d) in enumerate([("p11_d1", 64, 12, 20, 11, 1), ("p11_d2", 64, 12, 20, 11, 2),
                                                 ("p11_p7", 32, 3, 5, 11, 1), ("p5_d1", 16, 7, 9, 5, 1)]):
        g = torch.Generator().manual_seed(3000 + idx)
        x1 = torch.randn(2, C, H, W, generator=g)
        x2 = torch.randn(2, C, H, W, generator=g)
        y = t2s.correlate(x1, x2, patch_size=P, dilation_patch=d)
        out[f"{name}.x1"], out[f"{name}.x2"], out[f"{name}.y"] = x1.numpy(), x2.numpy(), y.numpy()
        out[f"{name}.pd"] = np.asarray([P, d], np.int32)
    # the concat + ReLU of TF_utils.py:28-31 on the first case
    x1, x2 = torch.from_numpy(out["p11_d1.x1"]), torch.from_numpy(out["p11_d1.x2"])
    g = torch.Generator().manual_seed(3999)
    t_ref, t_next = torch.randn(2, 24, 12, 20, generator=g), torch.randn(2, 24, 12, 20, generator=g)
    x_corr = t2s.correlate(x1, x2, patch_size=11)
    concat = torch.relu(torch.cat([x_corr, t_ref, t_next], dim=1))
    out["concat.t_ref"], out["concat.t_next"], out["concat.y"] = t_ref.numpy(), t_next.numpy(), concat.numpy()
    np.savez_compressed(os.path.join(OUT, "correlate.npz"), **out)


def _backbone_dcn():
    bb = rh.load_backbone()
    out = {}
    placement = {}
    for name, args in (("r50", ([3, 4, 6, 3], [0, 4, 6, 3], 2)), ("r101", ([3, 4, 23, 3], [0, 4, 23, 3], 3)),
                       ("r50_nodcn", ([3, 4, 6, 3],))):
        net = bb.ResNetBackbone(*args)
        placement[name] = [[li, bi] for li, layer in enumerate(net.layers) for bi, blk in enumerate(layer) if blk.use_dcn]
    out["placement"] = np.frombuffer(json.dumps(placement).encode(), dtype=np.uint8)
    # one stride-1 and one stride-2 Bottleneck with a DCN conv2 (reduced width to keep the fixture small)
    for name, stride in (("s1", 1), ("s2", 2)):
        torch.manual_seed(4000 + stride)
        blk = bb.Bottleneck(64, 16, stride=stride, use_dcn=True,
                            downsample=torch.nn.Sequential(torch.nn.Conv2d(64, 64, 1, stride=stride, bias=False),
                                                           torch.nn.BatchNorm2d(64)))
        blk.eval()
        dcn = blk.conv2
        torch.nn.init.normal_(dcn.conv_offset_mask.weight, std=0.05)     # non-zero
        torch.nn.init.normal_(dcn.conv_offset_mask.bias, std=0.5)
        torch.nn.init.normal_(dcn.bias, std=0.1)
        x = torch.randn(2, 64, 10, 14)
        cap = {}
        hk = dcn.register_forward_hook(lambda mod, inp, o, cap=cap: cap.update(x=inp[0].detach().clone(),
                                                                                y=o.detach().clone()))
        with torch.no_grad():
            yb = blk(x)
        hk.remove()
        out[f"{name}.dcn_x"], out[f"{name}.dcn_y"], out[f"{name}.block_y"] = cap["x"].numpy(), cap["y"].numpy(), yb.numpy()
        out[f"{name}.weight"], out[f"{name}.bias"] = dcn.weight.detach().numpy(), dcn.bias.detach().numpy()
        out[f"{name}.com_w"] = dcn.conv_offset_mask.weight.detach().numpy()
        out[f"{name}.com_b"] = dcn.conv_offset_mask.bias.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "backbone_dcn.npz"), **out)


def _detections():
    """One FPN level (24x40, one prior per pixel) through the reference's FeatureAlign head and Detect_TF."""
    fa_mod = rh.load_featurealign()
    det_mod = rh.load_detect()
    torch.manual_seed(4242)
    C, H, W, ks = 64, 24, 40, (3, 5)
    m = fa_mod.FeatureAlign(C, 41, kernel_size=ks, deformable_groups=1, use_pred_offset=True)
    torch.nn.init.normal_(m.conv_offset.weight, std=0.5)
    torch.nn.init.normal_(m.conv_adaption.weight, std=(C * 15) ** -0.5)
    torch.nn.init.normal_(m.conv.weight, std=4.0 * (C * 15) ** -0.5)       # confident logits, so NMS has work to do
    torch.nn.init.normal_(m.conv.bias, std=0.5)
    x, shape = torch.randn(1, C, H, W), torch.randn(1, 4, H, W)
    with torch.no_grad():
        logits = m(x.clone(), shape)                                           # [1, 41, H, W]
    conf = torch.softmax(logits[0].permute(1, 2, 0).reshape(H * W, 41), -1)    # one prior per pixel (STMask.py eval)
    # boxes: centred on the prior's pixel, random size, xyxy normalised; neighbours overlap heavily
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    cx, cy = ((xs + 0.5) / W).reshape(-1), ((ys + 0.5) / H).reshape(-1)
    bw, bh = 0.04 + 0.12 * torch.rand(H * W), 0.06 + 0.2 * torch.rand(H * W)
    boxes = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    centerness = 0.3 + 0.7 * torch.rand(H * W)
    # the reference's candidate filter (TF_utils.py:68-74, cfg.eval_conf_thresh = 0.05) ...
    conf_t = conf.t().contiguous()
    keep = torch.max(conf_t[1:, :], dim=0)[0] > 0.05
    cand = {"conf": conf_t[:, keep].t(), "box": boxes[keep], "centerness": centerness[keep],
            "mask_coeff": torch.zeros(int(keep.sum()), 8), "track": None, "proto": None}
    # ... and its cross-class fast NMS (detection_TF.py:56-134, nms_thresh 0.5, top_k 200)
    det = det_mod.Detect_TF(41, 0, 200, 0.05, 0.5).detect(cand)
    kept_prior = torch.nonzero(keep).view(-1)
    # identify the kept detections by their prior index (boxes are unique per prior)
    idx = [int(kept_prior[(cand["box"] == b).all(1).nonzero()[0, 0]]) for b in det["box"]]
    assert 10 <= len(idx) <= 200, len(idx)
    np.savez_compressed(os.path.join(OUT, "detections.npz"), x=x.numpy(), shape=shape.numpy(),
                        w_offset=m.conv_offset.weight.detach().numpy(), w_adaption=m.conv_adaption.weight.detach().numpy(),
                        w_conv=m.conv.weight.detach().numpy(), b_conv=m.conv.bias.detach().numpy(),
                        boxes=boxes.numpy(), centerness=centerness.numpy(), logits=logits.numpy(),
                        det_prior=np.asarray(idx, np.int64), det_class=det["class"].numpy().astype(np.int64),
                        det_score=det["score"].numpy())


def _roi_align():
    from torchvision.ops import roi_align as tv_roi_align
    out = {}
    g = torch.Generator().manual_seed(77)
    feat = torch.randn(2, 24, 12, 20, generator=g)
    rois = torch.tensor([[0, 1.5, 2.0, 9.3, 7.7], [1, -3.0, -2.0, 5.0, 4.0], [0, 10.0, 5.0, 25.0, 14.0], [1, 3.2, 3.1, 3.9, 3.6],
                         [0, 0.0, 0.0, 20.0, 12.0], [1, 18.5, 10.2, 19.9, 11.9], [0, 7.0, 3.0, 7.0, 3.0], [1, 30.0, 30.0, 40.0, 40.0]])
    out["tv.feat"], out["tv.rois"] = feat.numpy(), rois.numpy()
    cases = [("a7", (7, 7), 1.0, 0, True), ("a7_sr2", (7, 7), 1.0, 2, True), ("l7", (7, 7), 1.0, 0, False),
             ("a3x5_s05", (3, 5), 0.5, 0, True)]
    for name, osz, scale, sr, al in cases:
        out[f"tv.{name}.y"] = tv_roi_align(feat, rois, osz, scale, sr, al).numpy()
        out[f"tv.{name}.cfg"] = np.asarray([osz[0], osz[1], scale, sr, int(al)], np.float32)
    # the reference's call site: normalised boxes -> sanitize_coordinates_hw -> roi_align(feat, [0, x1, y1, x2, y2], 7)
    t2s = rh.load_track_to_segment_head()
    fmap = torch.randn(1, 48, 24, 40, generator=g)
    boxes = torch.rand(9, 4, generator=g)
    boxes[:, 2:] = (boxes[:, :2] + 0.05 + 0.4 * torch.rand(9, 2, generator=g))          # some cross the right / bottom border
    boxes[0] = torch.tensor([0.6, 0.7, 0.2, 0.1])                                          # swapped corners (sanitised)
    with torch.no_grad():
        pooled = t2s.bbox_feat_extractor(fmap, boxes.clone(), 24, 40, 7)
    out["ref.feat"], out["ref.boxes_norm"], out["ref.y"] = fmap.numpy(), boxes.numpy(), pooled.numpy()
    out["ref.boxes"] = rh.load_box_utils().sanitize_coordinates_hw(boxes.clone(), 24, 40).numpy()
    np.savez_compressed(os.path.join(OUT, "roi_align.npz"), **out)


TEMPORAL_NET_SEED = 20260117


def temporal_net_weight_checksums(net) -> np.ndarray:
    """[sum, sum of |.|] of every parameter in state_dict order: the GPU test rebuilds the weights from the seed
    (they are 40 MB — too big for a fixture) and must get exactly these."""
    return np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in net.state_dict().values()], np.float64)


def _temporal_net():
    """The reference's own TemporalNet.forward (track_to_segment_head.py:10-37) on 7x7 crops that its own
    bbox_feat_extractor (:65-88) cut out of a 633-channel concat (TF_utils.py:28-36)."""
    t2s = rh.load_track_to_segment_head()
    torch.manual_seed(TEMPORAL_NET_SEED)
    net = t2s.TemporalNet(633)            # default nn.Conv2d / nn.Linear init in construction order: conv1, conv2, conv3, fc, fc_coeff
    net.eval()
    g = torch.Generator().manual_seed(7)
    q = lambda t: t.bfloat16().float()    # exactly representable in bf16, so the bf16 device path sees the same inputs
    concat = torch.relu(q(torch.randn(2, 633, 12, 20, generator=g)))
    boxes = torch.rand(5, 4, generator=g)
    boxes = torch.stack([boxes[:, 0] * 0.6, boxes[:, 1] * 0.6, boxes[:, 0] * 0.6 + 0.1 + boxes[:, 2] * 0.3,
                         boxes[:, 1] * 0.6 + 0.1 + boxes[:, 3] * 0.3], 1)
    pair = torch.tensor([0, 0, 1, 1, 1])
    with torch.no_grad():
        crops = torch.cat([t2s.bbox_feat_extractor(concat[int(p)], boxes[i:i + 1], 12, 20, 7) for i, p in enumerate(pair)], 0)
        x_reg, x_coeff = net(crops)
        # intermediate activations, for localising a mismatch
        h1 = torch.relu(net.conv1(crops))
    np.savez_compressed(os.path.join(OUT, "temporal_net.npz"), seed=np.int64(TEMPORAL_NET_SEED),
                        checksums=temporal_net_weight_checksums(net), concat=concat.numpy().astype(np.float16) if False else concat.numpy(),
                        boxes=boxes.numpy(), pair=pair.numpy(), crops=crops.numpy(), x_reg=x_reg.numpy(), x_coeff=x_coeff.numpy(),
                        h1_mean=h1.mean(dim=(2, 3)).numpy())


HEAD_SEED = 20260303
HEAD_LEVELS = [(12, 20), (6, 10), (3, 5), (2, 3), (1, 2)]


def _prediction_head():
    """The reference's own PredictionModule_FC.forward (prediction_head_FC.py:129-222) of the R101 FCA+FCB(ada) config
    (`STMask_plus_base_ada_config`), applied to five FPN levels with shared weights and concatenated over the levels the
    way STMask.forward_single does (STMask.py:245-279).  Weights: default init under a recorded seed (7.6 M parameters:
    rebuilt, not stored; pinned by checksums)."""
    from . import ref_model
    ref_model._install_stubs()
    import STMask  # noqa: F401  (initialises the reference's package imports)
    from datasets.config import cfg, set_cfg
    set_cfg("STMask_plus_base_ada_config")
    import layers.modules.prediction_head_FC as ph
    cfg.mask_dim, cfg.num_heads = 32, 5
    torch.manual_seed(HEAD_SEED)
    head = ph.PredictionModule_FC(256, 256, deform_groups=1, pred_aspect_ratios=cfg.backbone.pred_aspect_ratios[0],
                                  pred_scales=cfg.backbone.pred_scales[0], parent=None)
    head.eval()
    g = torch.Generator().manual_seed(11)
    xs = [torch.randn(1, 256, h, w, generator=g).bfloat16().float() for h, w in HEAD_LEVELS]
    outs = {}
    with torch.no_grad():
        for x in xs:
            for k, v in head(x).items():
                outs.setdefault(k, []).append(v)
    res = {k: torch.cat(v, 1).numpy() for k, v in outs.items() if k != "T2S_feat"}
    res["T2S_feat0"] = outs["T2S_feat"][0].numpy()
    cs = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in head.state_dict().values()], np.float64)
    np.savez_compressed(os.path.join(OUT, "prediction_head.npz"), seed=np.int64(HEAD_SEED), checksums=cs,
                        **{f"x{i}": x.numpy() for i, x in enumerate(xs)}, **res)


def _mask_assembly():
    """The reference's own generate_mask (mask_utils.py:111-128: tanh coefficients, sigmoid, crop with 1 px padding) and
    mask_iou (box_utils.py:435-447) on two sets of detections over one prototype map."""
    from . import ref_model
    ref_model._install_stubs()
    import STMask  # noqa: F401
    from datasets.config import set_cfg
    set_cfg("STMask_plus_base_ada_config")
    import layers.mask_utils as mu
    import layers.box_utils as bu
    g = torch.Generator().manual_seed(21)
    proto = torch.relu(torch.randn(24, 40, 32, generator=g))
    out = {"proto": proto.numpy()}
    masks = []
    for name, n in (("a", 9), ("b", 6)):
        coeff = torch.randn(n, 32, generator=g)
        c = torch.rand(n, 2, generator=g)
        wh = 0.08 + torch.rand(n, 2, generator=g) * 0.4
        boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
        boxes[0] = torch.tensor([0.9, 0.2, 0.3, 0.7])            # x1 > x2: sanitize swaps them
        boxes[1] = torch.tensor([-0.2, -0.1, 1.3, 1.2])          # larger than the frame: clamped
        m = mu.generate_mask(proto, coeff, boxes)
        masks.append(m)
        out[f"{name}.coeff"], out[f"{name}.boxes"], out[f"{name}.masks"] = coeff.numpy(), boxes.numpy(), m.numpy()
    out["iou"] = bu.mask_iou(masks[0].gt(0.5).float(), masks[1].gt(0.5).float()).numpy()
    np.savez_compressed(os.path.join(OUT, "mask_assembly.npz"), **out)


TRACK_SEED = 77


def _tracker():
    """The reference's own Track_TF.track (track_TF.py:52-181) on a synthetic 7-frame clip: first frame, matches, two
    detections competing for one object, new objects, an empty frame (shift only), objects ageing, a video restart.
    `CandidateShift` is replaced by a recorded, seeded shift built from the reference's own decode / center_size /
    generate_mask (TF_utils.py:38-49 without the network): the matching state machine is what this fixture pins."""
    from . import ref_model
    ref_model._install_stubs()
    import STMask  # noqa: F401
    from datasets.config import cfg, set_cfg
    set_cfg("STMask_plus_base_ada_config")
    cfg.eval_conf_thresh = 0.3
    import layers.functions.track_TF as tt
    import layers.box_utils as bu
    import layers.mask_utils as mu
    assert cfg.train_track and list(cfg.match_coeff) == [0, 1, 2, 0]
    g = torch.Generator().manual_seed(TRACK_SEED)
    H, W, K, E = 24, 40, 32, 16
    n_ident = 9
    ident_emb = torch.nn.functional.normalize(torch.randn(n_ident, E, generator=g), dim=1)
    ident_c = 0.15 + 0.7 * torch.rand(n_ident, 2, generator=g)
    ident_wh = 0.12 + 0.2 * torch.rand(n_ident, 2, generator=g)
    ident_coeff = torch.randn(n_ident, K, generator=g)
    ident_cls = torch.randint(1, 41, (n_ident,), generator=g)
    # which identities are detected in each frame (an identity listed twice = two competing detections)
    frames = [([0, 1, 2, 3], True), ([1, 0, 4, 2, 2], False), ([], False), ([3, 5, 0, 0, 1], False), ([6, 2, 4], False),
              ([7, 8, 1], True), ([8, 7, 7, 0], False)]
    shifts, shifted = [], []

    def stub_shift(net, ref_candidate, next_candidate, img=None, img_meta=None, display=False):
        n = ref_candidate["box"].shape[0]
        loc = torch.randn(n, 4, generator=g) * 0.3
        dco = torch.randn(n, K, generator=g) * 0.05
        shifts.append((loc, dco))
        out = {k: v.clone() for k, v in next_candidate.items() if k in {"proto", "fpn_feat", "T2S_feat"}}
        box = bu.decode(loc, bu.center_size(ref_candidate["box"].clone()))
        coeff = ref_candidate["mask_coeff"].clone() + dco
        out["box"] = box.clone()
        out["score"] = ref_candidate["score"].clone() * 0.95
        out["mask_coeff"] = coeff.clone()
        out["mask"] = mu.generate_mask(next_candidate["proto"], coeff, box).clone()
        shifted.append({k: out[k].numpy().copy() for k in ("box", "score", "mask_coeff", "mask")})
        return out

    tt.CandidateShift = stub_shift
    tracker = tt.Track_TF()
    out = {"seed": np.int64(TRACK_SEED), "n_frames": np.int64(len(frames)), "conf_thresh": np.float32(cfg.eval_conf_thresh),
           "match_coeff": np.asarray(cfg.match_coeff, np.float32)}
    for f, (ids, first) in enumerate(frames):
        n = len(ids)
        idx = torch.tensor(ids, dtype=torch.long)
        proto = torch.relu(torch.randn(H, W, K, generator=g))
        c = ident_c[idx] + 0.01 * f + 0.01 * torch.randn(n, 2, generator=g)
        wh = ident_wh[idx] * (1 + 0.05 * torch.randn(n, 2, generator=g))
        cand = {"proto": proto, "T2S_feat": torch.zeros(1, 1, 2, 2), "fpn_feat": torch.zeros(1, 1, 2, 2),
                "box": torch.cat([c - wh / 2, c + wh / 2], 1),
                "score": 0.2 + 0.75 * torch.rand(n, generator=g),
                "class": ident_cls[idx].clone(),
                "mask_coeff": ident_coeff[idx] + 0.1 * torch.randn(n, K, generator=g),
                "track": torch.nn.functional.normalize(ident_emb[idx] + 0.15 * torch.randn(n, E, generator=g), dim=1),
                "centerness": torch.rand(n, generator=g)}
        if n == 0:
            cand["class"] = torch.zeros(0, dtype=torch.long)
        for k in ("box", "score", "class", "mask_coeff", "track", "centerness"):
            out[f"f{f}.det.{k}"] = cand[k].numpy().copy()
        out[f"f{f}.proto"] = proto.numpy()
        out[f"f{f}.is_first"] = np.bool_(first)
        n_shift = len(shifts)
        det = tracker.track(None, cand, {"is_first": first}, img=None)
        if len(shifts) > n_shift:
            out[f"f{f}.shift.loc"], out[f"f{f}.shift.coeff"] = shifts[-1][0].numpy(), shifts[-1][1].numpy()
            for k, v in shifted[-1].items():
                out[f"f{f}.shifted.{k}"] = v
        st = tracker.prev_candidate
        for k in ("box", "score", "class", "mask_coeff", "track", "centerness", "tracked_mask", "mask"):
            out[f"f{f}.state.{k}"] = st[k].numpy().copy()
        out[f"f{f}.out.box_ids"] = np.asarray(det["box_ids"].numpy() if det["box_ids"].numel() else np.zeros(0), np.int64)
        if det["box_ids"].numel():
            out[f"f{f}.out.box"] = det["box"].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "tracker.npz"), **out)


def _model_r50():
    """BASELINE.json configs[0] / SURVEY.md 8(c) "Model:" known answer: the reference's own STMask (R50-DCN-FPN
    FCA+TF) on a synthetic 2-frame clip through oracle/ref_model.py; every hot-op call site of frame 2."""
    from . import ref_model
    out = ref_model.run_two_frame_clip()
    keep = dict(out)
    keep["roi.out"] = out["roi.out"][:8]              # 8 of the 43 crops (5 MB otherwise); TemporalNet outputs are kept for all
    np.savez_compressed(os.path.join(OUT, "model_r50.npz"), **keep)


def main():
    if not rh.available():
        raise SystemExit("/root/reference is not present: fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # deterministic accumulation order
    _dcn_torchvision()
    _dcn_border()
    _feature_align()
    _correlate()
    _backbone_dcn()
    _detections()
    _roi_align()
    _temporal_net()
    _model_r50()
    _prediction_head()
    _mask_assembly()
    _tracker()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
