"""Temporal-fusion cost volume (reference layers/modules/track_to_segment_head.py:40-62 and the
concat + ReLU that consumes it, layers/functions/TF_utils.py:28-31, STMask.py:291-297)."""
from __future__ import annotations

import torch

from . import ops


def correlate(x1: torch.Tensor, x2: torch.Tensor, patch_size: int = 11, dilation_patch: int = 1) -> torch.Tensor:
    """Same signature and result as the reference's `correlate`: cost volume viewed as
    [B, P*P, H, W], divided by C, leaky-ReLU(0.1) — one kernel instead of four."""
    return ops.correlation(x1, x2, patch_size, dilation_patch, scale=1.0 / x1.size(1), leaky_slope=0.1)


def padded_corr_channels(patch_size: int = 11) -> int:
    """P*P rounded up to a multiple of 8 channels (121 -> 128): where the T2S features start in the padded layout."""
    return (patch_size * patch_size + 7) // 8 * 8


def correlate_concat(fpn_ref: torch.Tensor, fpn_next: torch.Tensor, t2s_ref: torch.Tensor, t2s_next: torch.Tensor,
                     patch_size: int = 11, dilation_patch: int = 1, channels_last: bool = True,
                     padded: bool = False) -> torch.Tensor:
    """relu(cat([correlate(fpn_ref, fpn_next), t2s_ref, t2s_next], dim=1)) in ONE kernel
    (TF_utils.py:30-31).  leaky-ReLU followed by ReLU is ReLU, so only the ReLU is applied.
    Returns the [B, P*P + 2*Ct, H, W] tensor that RoIAlign / TemporalNet consume.

    `padded=True` returns the B200 layout instead: [B, Cp + 2*Ct, H, W] channels-last with
    Cp = padded_corr_channels(P) (128 for P = 11): channels [0, P*P) = correlation, [P*P, Cp) = 0,
    [Cp, Cp+Ct) = relu(t2s_ref), [Cp+Ct, Cp+2Ct) = relu(t2s_next).  Every block of a pixel row is then
    16-byte aligned, so the kernel moves it with 16-byte loads/stores only; a consumer conv takes the
    layout by zero-padding its input-channel weights at [P*P, Cp) (`unpad_concat` gives the reference view)."""
    return ops.correlation(fpn_ref, fpn_next, patch_size, dilation_patch, scale=1.0 / fpn_ref.size(1), relu=True,
                           feats=(t2s_ref, t2s_next), channels_last=channels_last or padded,
                           feat_channel_offset=padded_corr_channels(patch_size) if padded else None)


def unpad_concat(x: torch.Tensor, patch_size: int = 11) -> torch.Tensor:
    """The reference's [B, P*P + 2*Ct, H, W] concat from the padded layout (a copy; for checks and for
    consumers that have not been re-laid-out)."""
    pp, cp = patch_size * patch_size, padded_corr_channels(patch_size)
    return torch.cat([x[:, :pp], x[:, cp:]], dim=1)


def pad_concat_weight(weight: torch.Tensor, patch_size: int = 11) -> torch.Tensor:
    """Input-channel re-layout of a conv weight that consumes the concat (TemporalNet.conv1,
    track_to_segment_head.py:10-37: [out, P*P + 2*Ct, kh, kw]) for the padded layout of `correlate_concat(padded=True)`:
    zero input channels are inserted at [P*P, Cp), so conv(x_padded, pad_concat_weight(w)) == conv(x_reference, w)
    exactly (the pad channels of x are zeros and meet zero weights).  Do it once, when the checkpoint is loaded."""
    pp, cp = patch_size * patch_size, padded_corr_channels(patch_size)
    if weight.dim() != 4 or weight.shape[1] <= pp:
        raise ValueError(f"expected a [out, {pp} + 2*Ct, kh, kw] weight, got {tuple(weight.shape)}")
    out = weight.new_zeros((weight.shape[0], weight.shape[1] + cp - pp) + tuple(weight.shape[2:]))
    out[:, :pp] = weight[:, :pp]
    out[:, cp:] = weight[:, pp:]
    return out
