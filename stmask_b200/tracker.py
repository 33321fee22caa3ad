"""Device-resident tracker — the host mirror of the reference's `Track_TF` (layers/functions/track_TF.py) for a BATCH
of independent clips, with no device->host synchronisation per frame.

The reference keeps `prev_candidate` as a dict of tensors that grows with `torch.cat`, decides every assignment in a
Python loop (`for idx, match_id in enumerate(match_ids)`, one sync per detection) and handles one video at a time.
Here the state is a set of fixed-capacity device tensors and a frame is a fixed sequence of launches of this library:

    shift (the caller's CandidateShift: correlation -> RoIAlign -> TemporalNet, temporal_net.shift_candidates)
      -> `apply_shift`: decode the shifted boxes, add the coefficient deltas, score x 0.95            (TF_utils.py:38-49)
      -> `stm_mask_assembly_fwd` for the shifted objects and for this frame's detections            (mask_utils.py:111-128)
      -> `stm_mask_iou_fwd`  detections x tracked objects on the bit planes                          (box_utils.py:435-447)
      -> `stm_track_update_fwd`  scores, arg-max, sequential assignment, row copies, output filter   (track_TF.py:104-165)

An object's id is its row in the state (`box_ids = arange`, track_TF.py:158); `keep` is the reference's output filter.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops

MATCH_COEFF = (0.0, 1.0, 2.0, 0.0)      # cfg.match_coeff of the STMask configs (datasets/config.py:685): score, mask IoU, box IoU, label


def decode_shift(loc: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """`decode(loc, center_size(box))` (box_utils.py:25-34,238-283; variances 0.1 / 0.2): the regressed shift applied to
    point-form boxes [..., 4] -> point-form boxes."""
    wh = boxes[..., 2:] - boxes[..., :2]
    ctr = (boxes[..., 2:] + boxes[..., :2]) / 2
    c = ctr + loc[..., :2] * 0.1 * wh
    s = wh * torch.exp(loc[..., 2:] * 0.2)
    x1y1 = c - s / 2
    return torch.cat([x1y1, x1y1 + s], -1)


class DeviceTracker:
    """State of `clips` independent videos, `cap` objects each (the reference's list is unbounded; objects past `cap`
    are dropped).  All tensors live on `device`; `step` never synchronises."""

    def __init__(self, clips: int, cap: int, mask_dim: int, embed_dim: int, proto_hw: Tuple[int, int], device,
                 match_coeff=MATCH_COEFF, conf_thresh: float = 0.05, max_age: int = 10, bbox_dummy_iou: float = 0.3):
        h, w = proto_hw
        words = (h * w + 31) // 32
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        self.state: Dict[str, torch.Tensor] = {
            "n_obj": z(clips, dt=torch.int32), "box": z(clips, cap, 4), "score": z(clips, cap), "cls": z(clips, cap, dt=torch.int32),
            "coeff": z(clips, cap, mask_dim), "track": z(clips, cap, embed_dim), "centerness": z(clips, cap),
            "tracked": z(clips, cap, dt=torch.int32), "mask_bits": z(clips, cap, words, dt=torch.int32), "mask": z(clips, cap, h, w)}
        self.clips, self.cap, self.proto_hw = clips, cap, (h, w)
        self.match_coeff, self.conf_thresh, self.max_age, self.bbox_dummy_iou = tuple(match_coeff), conf_thresh, max_age, bbox_dummy_iou

    @torch.no_grad()
    def apply_shift(self, loc_shift: torch.Tensor, coeff_shift: torch.Tensor, proto: torch.Tensor) -> None:
        """What `CandidateShift` writes back into `prev_candidate` (TF_utils.py:38-49): boxes decoded with the regressed
        shift, coefficients + their regressed delta, score x 0.95, masks regenerated on this frame's prototypes.
        loc_shift [clips, cap, 4], coeff_shift [clips, cap, k] (rows past n_obj are ignored), proto [clips, h, w, k]."""
        st = self.state
        st["box"].copy_(decode_shift(loc_shift.float(), st["box"]))
        st["coeff"].add_(coeff_shift.float())
        st["score"].mul_(0.95)
        self.refresh_masks(proto)

    @torch.no_grad()
    def refresh_masks(self, proto: torch.Tensor) -> None:
        st = self.state
        masks, bits = ops.mask_assembly(proto, st["coeff"], st["box"], st["n_obj"])
        live = (torch.arange(self.cap, device=masks.device)[None, :] < st["n_obj"][:, None])
        st["mask"].copy_(torch.where(live[:, :, None, None], masks, st["mask"]))
        st["mask_bits"].copy_(torch.where(live[:, :, None], bits, st["mask_bits"]))

    @torch.no_grad()
    def step(self, dets: Dict[str, torch.Tensor], proto: torch.Tensor, is_first: Optional[torch.Tensor] = None):
        """dets: count [clips] int32, box [clips, n, 4], score, cls (int32), coeff [clips, n, k], track [clips, n, e],
        centerness [clips, n] — this frame's detections after fast NMS (`ops.detect_fast_nms` + gathers); proto
        [clips, h, w, k].  The state must already carry CandidateShift's update (`apply_shift`).
        Returns (det_slot [clips, n], keep [clips, cap]); the tracked objects are `self.state` rows < n_obj."""
        d = dict(dets)
        d["mask"], d["mask_bits"] = ops.mask_assembly(proto, d["coeff"], d["box"], d.get("count"))
        # `is_first` clips start from an empty state inside the kernel; their stale IoU columns are never read
        miou = ops.mask_iou_bits(d["mask_bits"], self.state["mask_bits"], d.get("count"), self.state["n_obj"])
        return ops.track_update(self.state, d, miou, is_first, match_coeff=self.match_coeff, bbox_dummy_iou=self.bbox_dummy_iou,
                                conf_thresh=self.conf_thresh, max_age=self.max_age)
