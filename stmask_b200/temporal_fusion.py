"""Temporal-fusion cost volume (reference layers/modules/track_to_segment_head.py:40-62 and the
concat + ReLU that consumes it, layers/functions/TF_utils.py:28-31, STMask.py:291-297)."""
from __future__ import annotations

import torch

from . import ops


def correlate(x1: torch.Tensor, x2: torch.Tensor, patch_size: int = 11, dilation_patch: int = 1) -> torch.Tensor:
    """Same signature and result as the reference's `correlate`: cost volume viewed as
    [B, P*P, H, W], divided by C, leaky-ReLU(0.1) — one kernel instead of four."""
    return ops.correlation(x1, x2, patch_size, dilation_patch, scale=1.0 / x1.size(1), leaky_slope=0.1)


def correlate_concat(fpn_ref: torch.Tensor, fpn_next: torch.Tensor, t2s_ref: torch.Tensor, t2s_next: torch.Tensor,
                     patch_size: int = 11, dilation_patch: int = 1, channels_last: bool = True) -> torch.Tensor:
    """relu(cat([correlate(fpn_ref, fpn_next), t2s_ref, t2s_next], dim=1)) in ONE kernel
    (TF_utils.py:30-31).  leaky-ReLU followed by ReLU is ReLU, so only the ReLU is applied.
    Returns the [B, P*P + 2*Ct, H, W] tensor that RoIAlign / TemporalNet consume."""
    return ops.correlation(fpn_ref, fpn_next, patch_size, dilation_patch, scale=1.0 / fpn_ref.size(1), relu=True,
                           feats=(t2s_ref, t2s_next), channels_last=channels_last)
