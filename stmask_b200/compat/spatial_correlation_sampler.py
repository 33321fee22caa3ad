"""Drop-in for `spatial_correlation_sampler` as the reference uses it:
``from spatial_correlation_sampler import spatial_correlation_sample``
(reference track_to_segment_head.py:4, called at :53-59 with kernel_size=1, stride=1,
padding=0, patch_size=11, dilation_patch).  Output: (B, patch, patch, H, W)."""
from __future__ import annotations

import torch.nn as nn

from .. import ops

__all__ = ["spatial_correlation_sample", "SpatialCorrelationSampler"]


def _one(v, name):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise NotImplementedError(f"{name}={v}: only square values are supported")
        v = v[0]
    return int(v)


def spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
                               dilation_patch=1):
    k, s, p, d = _one(kernel_size, "kernel_size"), _one(stride, "stride"), _one(padding, "padding"), _one(dilation, "dilation")
    if (k, s, p, d) != (1, 1, 0, 1):
        raise NotImplementedError(
            f"kernel_size={k}, stride={s}, padding={p}, dilation={d}: the B200 kernels implement the configuration "
            f"STMask uses (kernel_size=1, stride=1, padding=0, dilation=1; any odd patch_size / dilation_patch)")
    P, dp = _one(patch_size, "patch_size"), _one(dilation_patch, "dilation_patch")
    out = ops.correlation(input1, input2, P, dp)
    b, _, h, w = out.shape
    return out.view(b, P, P, h, w)


class SpatialCorrelationSampler(nn.Module):
    def __init__(self, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
        super().__init__()
        self.kernel_size = kernel_size
        self.patch_size = patch_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.dilation_patch = dilation_patch

    def forward(self, input1, input2):
        return spatial_correlation_sample(input1, input2, self.kernel_size, self.patch_size, self.stride, self.padding,
                                          self.dilation, self.dilation_patch)
