"""Backbone DCN stages (reference backbone.py:8-58 `Bottleneck`, :105-138 `_make_layer`).

Only what the hot path needs: the placement rule that decides which bottlenecks carry a
DCNv2 `conv2`, the resulting layer shapes for a given frame size, and a `Bottleneck`-compatible
DCN branch built on the drop-in `DCN` module.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

from .compat.dcn_v2 import DCN

# (block counts, dcn_layers, dcn_interval) — reference datasets/config.py:285-289, 304-308
RESNET50_DCN = ([3, 4, 6, 3], [0, 4, 6, 3], 2)
RESNET101_DCN = ([3, 4, 23, 3], [0, 4, 23, 3], 3)


def dcn_block_flags(blocks: int, dcn_layers: int = 0, dcn_interval: int = 1) -> List[bool]:
    """use_dcn per bottleneck of one stage (reference backbone.py:124, :130)."""
    flags = [dcn_layers >= blocks]
    for i in range(1, blocks):
        flags.append(((i + dcn_layers) >= blocks) and (i % dcn_interval == 0))
    return flags


def dcn_placement(layers: Sequence[int], dcn_layers: Sequence[int] = (0, 0, 0, 0), dcn_interval: int = 1) -> List[Tuple[int, int]]:
    """(stage index, block index) of every DCN bottleneck, in forward order."""
    return [(li, bi) for li, (n, d) in enumerate(zip(layers, dcn_layers))
            for bi, f in enumerate(dcn_block_flags(n, d, dcn_interval)) if f]


@dataclass(frozen=True)
class DcnLayerShape:
    stage: int
    block: int
    channels: int     # planes: DCN is planes -> planes, 3x3, dg = 1
    in_h: int
    in_w: int
    stride: int
    out_h: int
    out_w: int

    @property
    def flops_per_frame(self) -> int:
        return 2 * self.out_h * self.out_w * self.channels * self.channels * 9


def dcn_layer_shapes(layers: Sequence[int], dcn_layers: Sequence[int], dcn_interval: int, height: int = 384,
                     width: int = 640) -> List[DcnLayerShape]:
    """Input/output geometry of every backbone DCN for a (padded) frame size.
    Stage s (0-based) works on planes 64*2^s; the stem reduces the frame by 4 (conv s2 + maxpool s2,
    backbone.py:90-93,143-146); stages 1..3 halve it in their first block (stride 2 in conv2)."""
    out = []
    h, w = (height + 1) // 2, (width + 1) // 2          # 7x7 s2 p3
    h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1           # maxpool 3x3 s2 p1
    for li, (n, d) in enumerate(zip(layers, dcn_layers)):
        planes = 64 * 2 ** li
        flags = dcn_block_flags(n, d, dcn_interval)
        for bi in range(n):
            stride = 2 if (bi == 0 and li > 0) else 1
            oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1     # 3x3, pad 1
            if flags[bi]:
                out.append(DcnLayerShape(li, bi, planes, h, w, stride, oh, ow))
            h, w = oh, ow
    return out


def make_bottleneck_dcn(planes: int, stride: int = 1, dilation: int = 1) -> DCN:
    """The DCN `conv2` of a bottleneck exactly as backbone.py:21-26 builds it (bias and the
    offset/mask predictor zero-initialised)."""
    conv2 = DCN(planes, planes, kernel_size=3, stride=stride, padding=dilation, dilation=dilation, deformable_groups=1)
    conv2.bias.data.zero_()
    conv2.conv_offset_mask.weight.data.zero_()
    conv2.conv_offset_mask.bias.data.zero_()
    return conv2
