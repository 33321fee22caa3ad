"""CPU restatement (numpy, plain loops) of the tracker's matching state machine — TEST INFRASTRUCTURE ONLY.

Follows the reference line by line:
  Track_TF.track          layers/functions/track_TF.py:52-181   (state handling, sequential assignment, output filter)
  compute_comp_scores     layers/functions/TF_utils.py:98-123   (dummy "new object" column, weighted sum left to right)
  jaccard / mask_iou      layers/box_utils.py:60-88, 435-447
`CandidateShift` (TF_utils.py:11-51) is NOT part of this step: the caller applies it to the state before calling
`track_update` (shifted boxes / mask coefficients / masks, score x 0.95).

Pinned by tests/golden/tracker.npz: the reference's own Track_TF.track run on a synthetic 6-frame clip
(oracle/make_golden.py::_tracker) — tests/test_oracle.py::test_track_oracle_reproduces_the_reference_tracker.

State / detections are dicts of numpy arrays: box [n,4], score [n], cls [n], coeff [n,k], track [n,e],
centerness [n], mask [n,h,w] (soft), and for the state tracked [n] ("tracked_mask").
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

KEYS = ("box", "score", "cls", "coeff", "track", "centerness", "mask")


def jaccard(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """box_utils.py:60-88 (point-form boxes, no clamping of the boxes)."""
    a = a.astype(np.float32)[:, None, :]
    b = b.astype(np.float32)[None, :, :]
    iw = np.maximum(np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]), 0).astype(np.float32)
    ih = np.maximum(np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]), 0).astype(np.float32)
    inter = iw * ih
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (area_a + area_b - inter)).astype(np.float32)


def mask_iou(m1: np.ndarray, m2: np.ndarray) -> np.ndarray:
    """box_utils.py:435-447 on {0,1} masks [n1,h,w], [n2,h,w]; 0 where the union is empty."""
    a = m1.reshape(m1.shape[0], -1).astype(np.float64)
    b = m2.reshape(m2.shape[0], -1).astype(np.float64)
    inter = a @ b.T
    union = a.sum(1)[:, None] + b.sum(1)[None, :] - inter
    out = np.zeros_like(inter)
    np.divide(inter, union, out=out, where=union != 0)
    return out.astype(np.float32)


def generate_mask(proto: np.ndarray, coeff: np.ndarray, boxes: np.ndarray, padding: int = 1) -> np.ndarray:
    """mask_utils.py:111-128 + crop / sanitize_coordinates (box_utils.py:296-316,341-364): sigmoid(proto . tanh(coeff))
    inside the box grown by `padding` pixels, 0 outside.  proto [h,w,k], coeff [n,k], boxes [n,4] relative -> [n,h,w]."""
    h, w, _ = proto.shape
    m = proto.astype(np.float32) @ np.tanh(coeff.astype(np.float32)).T                # [h, w, n]
    m = (1.0 / (1.0 + np.exp(-m.astype(np.float64)))).astype(np.float32)

    def sanitize(a, b, size):
        a, b = a.astype(np.float32) * size, b.astype(np.float32) * size
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        return np.maximum(lo - padding, 0), np.minimum(hi + padding, size)

    x1, x2 = sanitize(boxes[:, 0], boxes[:, 2], w)
    y1, y2 = sanitize(boxes[:, 1], boxes[:, 3], h)
    cols = np.arange(w, dtype=np.float32)[None, :, None]
    rows = np.arange(h, dtype=np.float32)[:, None, None]
    inside = (cols >= x1[None, None]) & (cols < x2[None, None]) & (rows >= y1[None, None]) & (rows < y2[None, None])
    return np.ascontiguousarray((m * inside).transpose(2, 0, 1), dtype=np.float32)


def decode_shift(loc: np.ndarray, boxes: np.ndarray) -> np.ndarray:
    """decode(loc, center_size(boxes)) (box_utils.py:25-34,238-283, variances 0.1 / 0.2) -> point-form boxes."""
    loc, boxes = loc.astype(np.float32), boxes.astype(np.float32)
    wh = boxes[:, 2:] - boxes[:, :2]
    ctr = (boxes[:, 2:] + boxes[:, :2]) / 2
    c = ctr + loc[:, :2] * np.float32(0.1) * wh
    s = wh * np.exp(loc[:, 2:] * np.float32(0.2))
    x1y1 = c - s / 2
    return np.concatenate([x1y1, x1y1 + s], 1).astype(np.float32)


def apply_shift(state: Dict[str, np.ndarray], loc: np.ndarray, dcoeff: np.ndarray, proto: np.ndarray) -> Dict[str, np.ndarray]:
    """What CandidateShift writes back into prev_candidate (TF_utils.py:38-49) given TemporalNet's outputs."""
    out = {k: v.copy() for k, v in state.items()}
    out["box"] = decode_shift(loc, state["box"])
    out["coeff"] = (state["coeff"] + dcoeff).astype(np.float32)
    out["score"] = (state["score"] * np.float32(0.95)).astype(np.float32)
    out["mask"] = generate_mask(proto, out["coeff"], out["box"])
    return out


def comp_scores(det: Dict[str, np.ndarray], prev: Dict[str, np.ndarray], match_coeff, bbox_dummy_iou: float = 0.3) -> np.ndarray:
    """track_TF.py:106-127 + TF_utils.py:98-123 -> [n_det, n_prev + 1] (column 0 = new object)."""
    n = det["box"].shape[0]
    cos = det["track"].astype(np.float32) @ prev["track"].astype(np.float32).T
    cos = np.concatenate([np.zeros((n, 1), np.float32), cos], 1)
    cos = (cos + 1) / 2
    dummy = np.full((n, 1), bbox_dummy_iou, np.float32)
    biou = np.concatenate([dummy, jaccard(det["box"], prev["box"])], 1)
    miou = np.concatenate([dummy, mask_iou(det["mask"] > 0.5, prev["mask"] > 0.5)], 1)
    label = np.concatenate([np.ones((n, 1), np.float32), (prev["cls"][None, :] == det["cls"][:, None]).astype(np.float32)], 1)
    c = [np.float32(v) for v in match_coeff]
    out = cos + c[0] * det["score"].astype(np.float32)[:, None]
    out = out + c[1] * miou
    out = out + c[2] * biou
    out = out + c[3] * label
    return out.astype(np.float32)


def track_update(state: Optional[Dict[str, np.ndarray]], det: Dict[str, np.ndarray], is_first: bool, match_coeff,
                 bbox_dummy_iou: float = 0.3, conf_thresh: float = 0.05, max_age: int = 10,
                 capacity: Optional[int] = None) -> Tuple[Optional[Dict[str, np.ndarray]], np.ndarray, np.ndarray]:
    """One frame of Track_TF.track after CandidateShift.  Returns (new state, det_slot [n_det], keep [n_obj]).
    `capacity`: the device kernel's fixed state size (new objects past it are dropped); None = unbounded like the reference."""
    if is_first:
        state = None
    n_det = det["box"].shape[0]
    det_slot = np.full(n_det, -1, np.int32)
    if n_det == 0 and state is None:
        return None, det_slot, np.zeros(0, bool)
    if n_det == 0:
        state = {k: v.copy() for k, v in state.items()}
        state["tracked"] = state["tracked"] + 1
    elif state is None:
        n0 = n_det if capacity is None else min(n_det, capacity)
        state = {k: det[k][:n0].copy() for k in KEYS}
        state["tracked"] = np.zeros(n0, np.int32)
        det_slot[:n0] = np.arange(n0)
    else:
        state = {k: v.copy() for k, v in state.items()}
        state["tracked"] = state["tracked"] + 1
        n_prev = state["box"].shape[0]
        comp = comp_scores(det, state, match_coeff, bbox_dummy_iou)
        match_ids = comp.argmax(1)                       # first maximum, like torch.max
        best_score = np.full(n_prev, -1.0, np.float32)
        best_idx = np.full(n_prev, -1, np.int64)
        for idx in range(n_det):
            m = int(match_ids[idx])
            if m == 0:
                if capacity is not None and state["box"].shape[0] >= capacity:
                    continue
                det_slot[idx] = state["box"].shape[0]
                for k in KEYS:
                    state[k] = np.concatenate([state[k], det[k][idx][None]], 0)
                state["tracked"] = np.concatenate([state["tracked"], np.zeros(1, np.int32)], 0)
            else:
                obj = m - 1
                if det["score"][idx] > best_score[obj]:
                    if best_idx[obj] != -1:
                        det_slot[best_idx[obj]] = -1
                    det_slot[idx] = obj
                    best_score[obj] = det["score"][idx]
                    best_idx[obj] = idx
                    for k in KEYS:
                        state[k][obj] = det[k][idx]
                    state["tracked"][obj] = 0
    keep = (state["tracked"] <= max_age) & ((state["mask"] > 0.5).reshape(state["mask"].shape[0], -1).sum(1) > 1) & \
           (state["score"] > np.float32(conf_thresh))
    return state, det_slot, keep
