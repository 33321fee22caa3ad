"""TemporalNet — the box / mask-coefficient regressor of the temporal-fusion module (reference
layers/modules/track_to_segment_head.py:10-37) and the `bbox_feat_extractor` in front of it (:65-88,
called from layers/functions/TF_utils.py:33-37).

Same constructor, parameter names and shapes as the reference (`conv1/2/3.{weight,bias}`, `fc`, `fc_coeff`), so a
released state_dict loads unchanged.  The B200 path keeps the 633-channel concat on the device in its padded
640-channel channels-last layout from the correlation kernel to the regressors:

    correlate_concat(padded=True)  ->  roi_align on the NHWC concat buffer  ->  conv1 (input channels re-laid out
    for the padding, ReLU fused)  ->  conv2  ->  conv3   [tcgen05 main loop of the deformable-conv kernel in its
    plain-conv mode: copy-only producers, UMMA, bias + ReLU epilogue]  ->  AvgPool(7x7) + fc + fc_coeff (one kernel)

No cuDNN anywhere on it.  fp32 inputs run the CUDA-core kernels of the same ABI (the <= 1e-4 path).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from . import ops
from .temporal_fusion import padded_corr_channels


class TemporalNet(nn.Module):
    def __init__(self, corr_channels, mask_proto_n=32, use_sipmask=False, sipmask_head=4, patch_size=11):
        super().__init__()
        self.conv1 = nn.Conv2d(corr_channels, 512, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(512, 512, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(512, 1024, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.pool = nn.AvgPool2d((7, 7), stride=1)
        self.fc = nn.Linear(1024, 4)
        self.fc_coeff = nn.Linear(1024, mask_proto_n * sipmask_head if use_sipmask else mask_proto_n)
        self.patch_size = patch_size
        self._convs = [ops.PlainConv() for _ in range(3)]
        self._fc_key = None
        self._fc = None

    def _fc_params(self):
        key = tuple((id(p), p._version, p.data_ptr()) for p in (self.fc.weight, self.fc.bias, self.fc_coeff.weight, self.fc_coeff.bias))
        if key != self._fc_key:
            self._fc = (torch.cat([self.fc.weight.detach().float(), self.fc_coeff.weight.detach().float()], 0).contiguous(),
                        torch.cat([self.fc.bias.detach().float(), self.fc_coeff.bias.detach().float()], 0).contiguous())
            self._fc_key = key
        return self._fc

    @torch.no_grad()
    def forward(self, x: torch.Tensor, padded: bool = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """x: [n, 633, 7, 7] (the reference's layout) or [n, 640, 7, 7] (the padded layout of
        `correlate_concat(padded=True)` after RoIAlign).  Returns (x_reg [n, 4], x_coeff [n, mask_proto_n]) in fp32."""
        pp = self.patch_size * self.patch_size
        cin = self.conv1.weight.shape[1]
        cp = padded_corr_channels(self.patch_size)
        if padded is None:
            padded = x.shape[1] == cin + cp - pp and x.shape[1] != cin
        if x.shape[1] != (cin + cp - pp if padded else cin):
            raise ValueError(f"expected {cin} (reference) or {cin + cp - pp} (padded) input channels, got {x.shape[1]}")
        in_pad = (pp, cp - pp) if padded else None
        h = self._convs[0]([x], self.conv1.weight, self.conv1.bias, 1, 1, relu=True, in_pad=in_pad)[0]
        h = self._convs[1]([h], self.conv2.weight, self.conv2.bias, 1, 1, relu=True)[0]
        h = self._convs[2]([h], self.conv3.weight, self.conv3.bias, 1, 1, relu=True)[0]
        w, b = self._fc_params()
        y = ops.pool_fc(h, w, b)
        return y[:, :4], y[:, 4:]


def sanitize_boxes_hw(boxes: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """Normalised xyxy boxes -> feature-map pixels, ordered and clamped exactly like the reference's
    `sanitize_coordinates_hw` (layers/box_utils.py:320-337, padding 0): x scaled by w, y by h, min/max swap, clamp."""
    b = boxes.float()
    x1, x2 = b[:, 0] * w, b[:, 2] * w
    y1, y2 = b[:, 1] * h, b[:, 3] * h
    xa, xb = torch.min(x1, x2).clamp(min=0), torch.max(x1, x2).clamp(max=w)
    ya, yb = torch.min(y1, y2).clamp(min=0), torch.max(y1, y2).clamp(max=h)
    return torch.stack([xa, ya, xb, yb], 1)


@torch.no_grad()
def shift_candidates(net: TemporalNet, concat: torch.Tensor, boxes_norm: torch.Tensor, pair_index: torch.Tensor,
                     pool_size: int = 7) -> Tuple[torch.Tensor, torch.Tensor]:
    """`bbox_feat_extractor` + `TemporalNet` for the boxes of MANY frame pairs at once (TF_utils.py:33-37 batched):
    concat [n_pairs, 640 | 633, H, W] straight from the correlation kernel, boxes_norm [n_boxes, 4] normalised xyxy
    of the reference frame's candidates, pair_index [n_boxes] = which pair each box belongs to.  The 7x7 crops, the
    three convs and the regressors all stay on the device in NHWC."""
    hh, ww = concat.shape[2:]
    rois = torch.cat([pair_index.float()[:, None], sanitize_boxes_hw(boxes_norm, hh, ww)], 1)
    crops = ops.roi_align(concat, rois, pool_size)
    return net(crops)
