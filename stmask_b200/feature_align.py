"""FCB — box-guided feature calibration (reference layers/modules/Featurealign.py:6-74).

Same constructor, parameter names (`conv_offset`, `conv_adaption`, `conv`) and forward
semantics as the reference's `FeatureAlign`, so a released state_dict loads unchanged.
B200-specific additions:
  * offsets come from one library kernel: `stm_fcb_ali_offsets` (closed form, instead of ~15
    elementwise torch ops) or `stm_fcb_ada_offsets` (the 1x1 conv_offset);
  * the ReLU after the deformable conv (Featurealign.py:72) is fused into its epilogue;
  * `forward_levels` runs ALL FPN levels of the weight-shared head (STMask.py:91-92,
    prediction_head_FC.py:157-167) in ONE grouped launch per conv — the output conv (Featurealign.py:73) included,
    on this library's own convolution kernel (no cuDNN).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .compat.mmcv_ops import DeformConv2d


class FeatureAlign(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), deformable_groups=4, use_pred_offset=True):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size, kernel_size)
        self.kernel_size = tuple(kernel_size)
        self.padding = ((kernel_size[0] - 1) // 2, (kernel_size[1] - 1) // 2)
        self.use_pred_offset = use_pred_offset
        self.deformable_groups = deformable_groups
        if self.use_pred_offset:
            offset_channels = kernel_size[0] * kernel_size[1] * 2
            self.conv_offset = nn.Conv2d(4, deformable_groups * offset_channels, 1, bias=False)
        self.conv_adaption = DeformConv2d(in_channels, in_channels, kernel_size=self.kernel_size, padding=self.padding,
                                          deform_groups=deformable_groups)
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=self.kernel_size, padding=self.padding)
        self._fused = None          # None: untried, True / False: the fused-offset tcgen05 path applies / does not
        self._out_conv = ops.PlainConv()

    def init_weights(self, bias_value=0):
        if self.use_pred_offset:
            torch.nn.init.normal_(self.conv_offset.weight, std=0.0)
        torch.nn.init.normal_(self.conv_adaption.weight, std=0.01)

    def offsets(self, shape: torch.Tensor) -> torch.Tensor:
        """ada: 1x1 conv of the (detached) box deltas; ali: closed form (Featurealign.py:43-69)."""
        if self.use_pred_offset:
            return ops.fcb_ada_offsets(shape.detach(), self.conv_offset.weight)
        if self.deformable_groups != 1:
            raise ValueError("FCB(ali) offsets are defined for deformable_groups == 1 (Featurealign.py:67-69)")
        return ops.fcb_ali_offsets(shape.detach(), self.kernel_size)

    def calibrate_levels(self, xs: Sequence[torch.Tensor], shapes: Sequence[torch.Tensor],
                         outs: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
        """relu(conv_adaption(x, offset)) for every level, one launch."""
        ops._no_grad_inputs(self.conv_adaption.weight, *xs)
        spec = self.conv_adaption.spec()
        wp = self.conv_adaption._cache.weight(self.conv_adaption.weight, spec, xs[0].dtype)
        if self._fused is not False and xs[0].dtype == torch.bfloat16 and (self.use_pred_offset or self.deformable_groups == 1):
            # tcgen05 path: the offsets are derived inside the sampling kernel from the box deltas — no offset tensors,
            # no offset kernels.  Shapes that need the CUDA-core kernel answer STM_ERR_UNSUPPORTED once; remember it.
            try:
                ys = ops.deform_conv2d_fcb_multi(list(xs), [s.detach() for s in shapes], wp, spec,
                                                 self.conv_offset.weight if self.use_pred_offset else None, relu=True, outs=outs)
                self._fused = True
                return ys
            except ops.L.StmError as e:
                if "status -2" not in str(e):
                    raise
                if self._fused is None:
                    self._fused = False
        offs = [self.offsets(s) for s in shapes]
        return ops.deform_conv2d_multi(list(xs), offs, None, wp, None, spec, relu=True, outs=outs)

    def forward_levels(self, xs: Sequence[torch.Tensor], shapes: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """conv(relu(conv_adaption(x, offset))) for every level (Featurealign.py:72-73): two launches for all levels, both this
        library's (the output conv on the TMA shifted-view kernel for bf16 activations, on the CUDA-core kernel for fp32)."""
        cal = self.calibrate_levels(xs, shapes)
        c = self.conv
        ys = self._out_conv(cal, c.weight, c.bias, c.stride, c.padding, c.dilation)
        return [y[:, :c.out_channels] for y in ys]

    def forward(self, x, shape):
        return self.forward_levels([x], [shape])[0]
