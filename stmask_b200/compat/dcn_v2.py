"""Drop-in for `dcn_v2` (CharlesShang/DCNv2) as the reference uses it:
``from dcn_v2 import DCN`` (reference backbone.py:5, FPN.py:8, FastMaskIoUNet.py:8),
constructed at backbone.py:21-22 and called at backbone.py:45.

Same constructor arguments, parameter names and shapes (`weight`, `bias`,
`conv_offset_mask.{weight,bias}`) so released checkpoints load unchanged."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops

__all__ = ["DCNv2", "DCN", "dcn_v2_conv"]


def dcn_v2_conv(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
    """Functional DCNv2: `mask` is already sigmoid-ed, as in the original autograd function."""
    return ops.deform_conv2d(input, offset, weight, bias, mask, stride, padding, dilation, 1, deformable_groups)


class DCNv2(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = ops._pair(kernel_size)
        self.stride = ops._pair(stride)
        self.padding = ops._pair(padding)
        self.dilation = ops._pair(dilation)
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()
        self._cache = ops.PackedWeightCache()

    def reset_parameters(self):
        n = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        stdv = 1.0 / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        self.bias.data.zero_()

    def _spec(self):
        return ops.ConvSpec(self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding,
                            self.dilation, 1, self.deformable_groups)

    def forward(self, input, offset, mask):
        k = self.kernel_size[0] * self.kernel_size[1]
        if offset.shape[1] != 2 * self.deformable_groups * k or mask.shape[1] != self.deformable_groups * k:
            raise ValueError("offset/mask channel count does not match kernel_size and deformable_groups")
        ops._no_grad_inputs(self.weight, self.bias)
        spec = self._spec()
        return ops.deform_conv2d_multi([input], [offset], [mask], self._cache.weight(self.weight, spec, input.dtype),
                                       self._cache.bias(self.bias), spec)[0]


class DCN(DCNv2):
    """DCNv2 with its own offset/mask predictor (a regular conv with 3*dg*kh*kw channels)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
        channels_ = self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(self.in_channels, channels_, kernel_size=self.kernel_size,
                                          stride=self.stride, padding=self.padding, bias=True)
        self.init_offset()
        self._predictor = ops.PlainConv()

    def init_offset(self):
        self.conv_offset_mask.weight.data.zero_()
        self.conv_offset_mask.bias.data.zero_()

    def forward(self, input, out=None):
        # The original does conv_offset_mask -> chunk -> cat(o1, o2) -> sigmoid(mask) -> dcn_v2_conv.  Here the
        # predictor is this library's own regular convolution (zero-offset mode of the same tcgen05 main loop)
        # with the bias fused and the fp32 accumulators stored as fp32: sampling positions are never rounded to
        # bf16.  cat(o1, o2) is the first 2/3 of its channels, so the sampling kernel reads offsets and mask logits
        # straight out of that tensor (views, no copies) and applies the sigmoid while sampling.  On the tcgen05 path the
        # predictor writes it plane-major ([B, 32, Ho, Wo] contiguous): the sampling kernel's per-tap loads of 32
        # neighbouring pixels' offsets are then one 128-byte line instead of 32.
        ops._no_grad_inputs(input, self.weight, self.bias, self.conv_offset_mask.weight, self.conv_offset_mask.bias)
        n_off = 2 * self.deformable_groups * self.kernel_size[0] * self.kernel_size[1]
        n_all = n_off + n_off // 2
        com = self.conv_offset_mask
        om = self._predictor([input], com.weight, com.bias, com.stride, com.padding, com.dilation, out_f32=True, out_planar=True)[0]
        spec = self._spec()
        return ops.deform_conv2d_multi([input], [om[:, :n_off]], [om[:, n_off:n_all]],
                                       self._cache.weight(self.weight, spec, input.dtype),
                                       self._cache.bias(self.bias), spec, mask_sigmoid=True,
                                       outs=None if out is None else [out])[0]
