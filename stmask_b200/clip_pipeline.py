"""The test-time tail of STMask for a BATCH of clips, device-resident end to end: what `STMask.forward` does after the
prediction heads (reference STMask.py:205-282 -> layers/functions/detection_TF.py -> layers/functions/track_TF.py ->
layers/functions/TF_utils.py), composed from this library's kernels with no device->host synchronisation per frame:

    head outputs of frame t (loc / conf / centerness / mask_coeff / track / priors, e.g. PredictionHeadFC.forward_levels)
      -> softmax, `stm_detect_fast_nms_fwd`                      generate_candidate + Detect_TF.cc_fast_nms, all clips at once
      -> gathers of the survivors' coefficients / embeddings     (index tensors stay on the device)
      -> CandidateShift for every tracked object of every clip:  correlation + concat (`correlate_concat`, padded layout)
         -> RoIAlign -> TemporalNet (`shift_candidates`) -> decode / coefficient delta / score x 0.95 / masks (`apply_shift`)
      -> mask assembly, bit-plane mask IoU, `stm_track_update_fwd`   Track_TF.track's matching state machine

The reference handles one video at a time with a Python loop and a sync per detection; here a "frame" is a fixed sequence
of launches whatever the clips contain.  Rows past `n_obj` / `count` are computed and ignored.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops
from .temporal_fusion import correlate_concat
from .temporal_net import TemporalNet, shift_candidates
from .tracker import MATCH_COEFF, DeviceTracker


class ClipPipeline:
    def __init__(self, temporal_net: TemporalNet, clips: int, cap: int, mask_dim: int, embed_dim: int, proto_hw: Tuple[int, int], device,
                 top_k: int = 100, conf_thresh: float = 0.05, nms_thresh: float = 0.5, match_coeff=MATCH_COEFF, max_age: int = 10):
        self.net = temporal_net
        self.tracker = DeviceTracker(clips, cap, mask_dim, embed_dim, proto_hw, device, match_coeff=match_coeff,
                                     conf_thresh=conf_thresh, max_age=max_age)
        self.clips, self.cap, self.top_k = clips, cap, top_k
        self.conf_thresh, self.nms_thresh = conf_thresh, nms_thresh
        self._prev: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._pair_index = torch.arange(clips, device=device).repeat_interleave(cap)

    @torch.no_grad()
    def detect(self, preds: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """generate_candidate + cross-class fast NMS (TF_utils.py:54-82, detection_TF.py:85-134) for one frame of every clip
        -> padded detection tensors [clips, top_k, ...] + count [clips]."""
        conf = torch.softmax(preds["conf"].float(), -1)
        count, index, cls, score, box = ops.detect_fast_nms(conf, preds["loc"], preds.get("centerness"), preds["priors"],
                                                            conf_thresh=self.conf_thresh, nms_thresh=self.nms_thresh, top_k=self.top_k)
        idx = index.clamp(min=0).long()
        take = lambda t: torch.gather(t.float(), 1, idx[..., None].expand(-1, -1, t.shape[-1])).contiguous()
        dets = {"count": count, "index": index, "box": box, "score": score, "cls": cls,
                "coeff": take(preds["mask_coeff"]), "track": take(preds["track"])}
        ctr = preds.get("centerness")
        dets["centerness"] = take(ctr.reshape(ctr.shape[0], -1, 1))[..., 0].contiguous() if ctr is not None else torch.zeros_like(score)
        return dets

    @torch.no_grad()
    def step(self, preds: Dict[str, torch.Tensor], fpn_feat: torch.Tensor, t2s_feat: torch.Tensor, proto: torch.Tensor,
             is_first: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """One frame of every clip.  fpn_feat / t2s_feat [clips, 256, H, W]: the level the reference correlates (P3) and the head's
        T2S features of it; proto [clips, h, w, k].  Returns the detections, the regressed shifts and (det_slot, keep); the
        tracked objects are `self.tracker.state` rows < n_obj."""
        out = self.detect(preds)
        if self._prev is not None:
            x640 = correlate_concat(self._prev[0], fpn_feat, self._prev[1], t2s_feat, channels_last=True, padded=True)
            boxes = self.tracker.state["box"].reshape(-1, 4)
            loc, dco = shift_candidates(self.net, x640, boxes, self._pair_index)
            out["loc_shift"], out["coeff_shift"] = loc.reshape(self.clips, self.cap, 4), dco.reshape(self.clips, self.cap, -1)
            self.tracker.apply_shift(out["loc_shift"], out["coeff_shift"], proto)
        else:
            self.tracker.refresh_masks(proto)
        dets = {k: out[k] for k in ("count", "box", "score", "cls", "coeff", "track", "centerness")}
        out["det_slot"], out["keep"] = self.tracker.step(dets, proto, is_first)
        self._prev = (fpn_feat, t2s_feat)
        return out
