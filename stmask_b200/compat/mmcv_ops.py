"""Drop-in for the part of `mmcv.ops` (mmcv-full 1.1.2) on STMask's hot path:
``from mmcv.ops import DeformConv2d, roi_align`` (reference Featurealign.py:3,
prediction_head_FC.py:10, track_to_segment_head.py:6) plus the modulated variants the
north star names.  Standard (padH, padW) semantics, i.e. what the reference's manual patch
of mmcv/ops/deform_conv.py restores (reference README.md:63-88).  Unlike mmcv there is no
`im2col_step` batch-divisibility limit (mmcv asserts B % min(32, B) == 0)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops

__all__ = ["DeformConv2d", "DeformConv2dPack", "ModulatedDeformConv2d", "ModulatedDeformConv2dPack",
           "deform_conv2d", "modulated_deform_conv2d", "roi_align"]


def deform_conv2d(input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deform_groups=1, bias=False,
                  im2col_step=32):
    if bias:
        raise AssertionError("Only support bias is False.")
    if input.dim() != 4:
        raise ValueError(f"Expected 4D tensor as input, got {input.dim()}D tensor instead.")
    return ops.deform_conv2d(input, offset, weight, None, None, stride, padding, dilation, groups, deform_groups)


def modulated_deform_conv2d(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                            deform_groups=1):
    if input.dim() != 4:
        raise ValueError(f"Expected 4D tensor as input, got {input.dim()}D tensor instead.")
    return ops.deform_conv2d(input, offset, weight, bias, mask, stride, padding, dilation, groups, deform_groups)


class DeformConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=False):
        super().__init__()
        assert not bias, f"bias={bias} is not supported in DeformConv2d."
        assert in_channels % groups == 0, f"in_channels {in_channels} cannot be divisible by groups {groups}"
        assert out_channels % groups == 0, f"out_channels {out_channels} cannot be divisible by groups {groups}"
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = ops._pair(kernel_size)
        self.stride = ops._pair(stride)
        self.padding = ops._pair(padding)
        self.dilation = ops._pair(dilation)
        self.groups = groups
        self.deform_groups = deform_groups
        self.transposed = False
        self.output_padding = (0, 0)
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // self.groups, *self.kernel_size))
        self.reset_parameters()
        self._cache = ops.PackedWeightCache()

    def reset_parameters(self):
        n = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)

    def spec(self):
        return ops.ConvSpec(self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding,
                            self.dilation, self.groups, self.deform_groups)

    def forward(self, x, offset, relu=False):
        # mmcv zero-pads inputs smaller than the kernel and crops the result; sampling outside the map
        # already reads zeros here, so the result is identical without the extra copies.
        if x.dim() != 4:
            raise ValueError(f"Expected 4D tensor as input, got {x.dim()}D tensor instead.")
        ops._no_grad_inputs(self.weight)
        spec = self.spec()
        return ops.deform_conv2d_multi([x], [offset], None, self._cache.weight(self.weight, spec, x.dtype), None, spec,
                                       relu=relu)[0]


class DeformConv2dPack(DeformConv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deform_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=self.stride, padding=self.padding,
                                     dilation=self.dilation, bias=True)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()
        self._predictor = ops.PlainConv()

    def forward(self, x):
        # the offset predictor is this library's own regular convolution (fp32 output: sampling positions are never rounded
        # to bf16; plane-major so that the sampling kernel's per-tap offset loads coalesce), like dcn_v2.DCN's
        co = self.conv_offset
        n_off = co.weight.shape[0]
        off = self._predictor([x], co.weight, co.bias, co.stride, co.padding, co.dilation, out_f32=True, out_planar=True)[0]
        return super().forward(x, off[:, :n_off])


class ModulatedDeformConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = ops._pair(kernel_size)
        self.stride = ops._pair(stride)
        self.padding = ops._pair(padding)
        self.dilation = ops._pair(dilation)
        self.groups = groups
        self.deform_groups = deform_groups
        self.transposed = False
        self.output_padding = (0, 0)
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.init_weights()
        self._cache = ops.PackedWeightCache()

    def init_weights(self):
        n = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def spec(self):
        return ops.ConvSpec(self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding,
                            self.dilation, self.groups, self.deform_groups)

    def forward(self, x, offset, mask, mask_sigmoid=False):
        ops._no_grad_inputs(self.weight, self.bias)
        spec = self.spec()
        return ops.deform_conv2d_multi([x], [offset], [mask], self._cache.weight(self.weight, spec, x.dtype),
                                       self._cache.bias(self.bias), spec, mask_sigmoid=mask_sigmoid)[0]


class ModulatedDeformConv2dPack(ModulatedDeformConv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(self.in_channels, self.deform_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
                                     kernel_size=self.kernel_size, stride=self.stride, padding=self.padding,
                                     dilation=self.dilation, bias=True)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()
        self._predictor = ops.PlainConv()

    def forward(self, x):
        co = self.conv_offset
        out = self._predictor([x], co.weight, co.bias, co.stride, co.padding, co.dilation, out_f32=True, out_planar=True)[0]
        n_off = 2 * self.deform_groups * self.kernel_size[0] * self.kernel_size[1]
        return super().forward(x, out[:, :n_off], out[:, n_off:n_off + n_off // 2], mask_sigmoid=True)


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True):
    """mmcv.ops.roi_align as `bbox_feat_extractor` calls it (reference track_to_segment_head.py:85-86):
    rois [n, 5] = (batch index, x1, y1, x2, y2) in feature-map pixels, average pooling.  Runs this library's
    NHWC kernel (`stm_roi_align_fwd`); CUDA only, like every operator here."""
    if pool_mode != "avg":
        raise NotImplementedError("only pool_mode='avg' is available")
    return ops.roi_align(input, rois, output_size, spatial_scale, sampling_ratio, aligned)
