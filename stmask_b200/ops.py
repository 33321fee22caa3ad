"""Functional front-end: torch tensors -> the C ABI of libstmask_b200.so.

PyTorch is plumbing here (device memory, streams); the arithmetic is in the CUDA library.
Every function raises on CPU tensors — there is no CPU implementation and no eager fallback.

Tensor conventions: NCHW-*shaped* tensors in any memory format.  Kernels read
NHWC, so channels-last tensors are consumed in place; contiguous-NCHW tensors are converted
once by the library's own layout kernel.  Outputs are channels-last tensors of NCHW shape.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _lib as L

IntPair = Union[int, Sequence[int]]

_BACKENDS = {"auto": L.BACKEND_AUTO, "simt": L.BACKEND_SIMT, "tcgen05": L.BACKEND_TCGEN05}


def _pair(v: IntPair) -> Tuple[int, int]:
    if isinstance(v, int):
        return v, v
    v = tuple(int(i) for i in v)
    if len(v) == 1:
        return v[0], v[0]
    if len(v) != 2:
        raise ValueError(f"expected an int or a pair, got {v}")
    return v


def _dt(t: torch.Tensor, what: str) -> int:
    if t.dtype == torch.float32:
        return L.STM_F32
    if t.dtype == torch.bfloat16:
        return L.STM_BF16
    raise TypeError(f"{what}: dtype {t.dtype} not supported (float32 or bfloat16)")


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: stmask_b200 operators run on CUDA (sm_100a) only; "
                           f"there is no CPU implementation")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def out_size(n: int, k: int, s: int, p: int, d: int) -> int:
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Return a tensor of the same NCHW shape whose channel stride is 1 (NHWC memory)."""
    if x.dim() != 4:
        raise ValueError(f"expected a 4-D NCHW tensor, got {tuple(x.shape)}")
    n, c, h, w = x.shape
    if x.numel() > 0 and (c == 1 or x.stride(1) == 1) and x.stride(3) >= c and x.stride(2) >= 0 and x.stride(0) >= 0:
        return x            # channel stride 1: the kernels take the other three strides explicitly
    if x.numel() == 0:
        return x.contiguous(memory_format=torch.channels_last)
    if x.is_contiguous() and n <= 65535:
        y = torch.empty_like(x, memory_format=torch.channels_last)
        dt = _dt(x, "x")
        with torch.cuda.device(x.device):
            L.check(L.lib().stm_nchw_to_nhwc(x.data_ptr(), dt, y.data_ptr(), dt, n, c, h, w, _stream(x)), "stm_nchw_to_nhwc")
        return y
    return x.contiguous(memory_format=torch.channels_last)


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    """Contiguous-NCHW copy of a channels-last tensor (library layout kernel)."""
    if x.is_contiguous():
        return x
    n, c, h, w = x.shape
    if x.is_contiguous(memory_format=torch.channels_last) and n <= 65535 and x.numel() > 0:
        y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        dt = _dt(x, "x")
        with torch.cuda.device(x.device):
            L.check(L.lib().stm_nhwc_to_nchw(x.data_ptr(), dt, y.data_ptr(), dt, n, c, h, w, _stream(x)), "stm_nhwc_to_nchw")
        return y
    return x.contiguous()


# --------------------------------------------------------------------------------------------
# deformable convolution
# --------------------------------------------------------------------------------------------
class ConvSpec:
    """Static parameters of one deformable conv (shared by all problems of a launch)."""

    __slots__ = ("in_c", "out_c", "kernel", "stride", "padding", "dilation", "groups", "deform_groups")

    def __init__(self, in_c: int, out_c: int, kernel: IntPair, stride: IntPair = 1, padding: IntPair = 0,
                 dilation: IntPair = 1, groups: int = 1, deform_groups: int = 1):
        self.in_c, self.out_c = int(in_c), int(out_c)
        self.kernel, self.stride = _pair(kernel), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.groups, self.deform_groups = int(groups), int(deform_groups)
        if self.in_c % self.groups or self.out_c % self.groups:
            raise ValueError(f"in_channels {in_c} / out_channels {out_c} must be divisible by groups {groups}")
        if self.in_c % self.deform_groups:
            raise ValueError(f"in_channels {in_c} must be divisible by deform_groups {deform_groups}")

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        return (out_size(h, self.kernel[0], self.stride[0], self.padding[0], self.dilation[0]),
                out_size(w, self.kernel[1], self.stride[1], self.padding[1], self.dilation[1]))

    def c_struct(self, dtype: int, offset_dtype: int, flags: int, backend: int) -> L.StmDcnConv:
        return L.StmDcnConv(self.in_c, self.out_c, self.kernel[0], self.kernel[1], self.stride[0], self.stride[1],
                            self.padding[0], self.padding[1], self.dilation[0], self.dilation[1], self.groups,
                            self.deform_groups, dtype, offset_dtype, flags, backend)


def pack_weight(weight: torch.Tensor, spec: ConvSpec, dtype: torch.dtype) -> torch.Tensor:
    """[Co, Ci/g, kh, kw] -> OHWI in `dtype` (one library launch)."""
    _require_cuda(weight, "weight")
    exp = (spec.out_c, spec.in_c // spec.groups, spec.kernel[0], spec.kernel[1])
    if tuple(weight.shape) != exp:
        raise ValueError(f"weight shape {tuple(weight.shape)} != {exp}")
    src = weight.detach()
    if src.dtype not in (torch.float32, torch.bfloat16):
        src = src.float()
    src = src.contiguous()
    packed = torch.empty((spec.out_c, spec.kernel[0], spec.kernel[1], spec.in_c // spec.groups), dtype=dtype, device=weight.device)
    conv = spec.c_struct(_dt(packed, "packed weight"), L.STM_F32, 0, L.BACKEND_AUTO)
    with torch.cuda.device(weight.device):
        L.check(L.lib().stm_dcn_pack_weight(C.byref(conv), src.data_ptr(), _dt(src, "weight"), packed.data_ptr(),
                                            _stream(weight)), "stm_dcn_pack_weight")
    return packed


class PackedWeightCache:
    """Packed weights (and fp32 biases) of a module, re-packed only after the source tensor was modified.

    An entry remembers the source tensor by WEAK REFERENCE together with its version counter and storage
    address: a hit needs the very same tensor object (an address recycled by the caching allocator after the
    original was freed can never alias), the same `_version` (in-place ops, optimizer steps, `load_state_dict`)
    and the same `data_ptr()` / dtype (`module.to(...)`, `param.data = ...`).  The one thing this cannot see is
    an in-place write THROUGH `.data` (`w.data.copy_(...)`, EMA swaps): autograd hides those from the version
    counter by design — call `invalidate()` (or `module._cache.invalidate()`) after such a write."""

    def __init__(self, max_entries: int = 256):
        self._w = {}
        self._b = {}
        self._max = max_entries

    def invalidate(self) -> None:
        self._w.clear()
        self._b.clear()

    @staticmethod
    def _alive(entry, t: torch.Tensor) -> bool:
        return entry is not None and entry[0]() is t and entry[1] == (t._version, t.data_ptr(), t.dtype)

    def weight(self, weight: torch.Tensor, spec: ConvSpec, dtype: torch.dtype) -> torch.Tensor:
        import weakref
        key = (id(weight), dtype)
        hit = self._w.get(key)
        if not self._alive(hit, weight):
            if len(self._w) >= self._max:
                self._w.clear()
            hit = (weakref.ref(weight), (weight._version, weight.data_ptr(), weight.dtype), pack_weight(weight, spec, dtype))
            self._w[key] = hit
        return hit[2]

    def bias(self, bias: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if bias is None:
            return None
        import weakref
        key = id(bias)
        hit = self._b.get(key)
        if not self._alive(hit, bias):
            if len(self._b) >= self._max:
                self._b.clear()
            hit = (weakref.ref(bias), (bias._version, bias.data_ptr(), bias.dtype), bias.detach().float().contiguous())
            self._b[key] = hit
        return hit[2]


def _no_grad_inputs(*tensors: Optional[torch.Tensor]) -> None:
    """The library is forward-only (the north star's scope): refuse, loudly, to produce outputs without a
    grad_fn for inputs that require grad instead of silently cutting the autograd graph."""
    if torch.is_grad_enabled():
        for t in tensors:
            if t is not None and t.requires_grad:
                raise RuntimeError("stmask_b200 operators are forward-only (no autograd): call them under torch.no_grad() "
                                   "or detach the inputs; an input or weight requires grad")


def _check_offset(t: Optional[torch.Tensor], ch: int, b: int, ho: int, wo: int, what: str) -> None:
    if t is None:
        return
    _require_cuda(t, what)
    if tuple(t.shape) != (b, ch, ho, wo):
        raise ValueError(f"{what} shape {tuple(t.shape)} != {(b, ch, ho, wo)}")


def deform_conv2d_multi(xs: Sequence[torch.Tensor], offsets: Sequence[Optional[torch.Tensor]],
                        masks: Optional[Sequence[Optional[torch.Tensor]]], w_packed: torch.Tensor,
                        bias_f32: Optional[torch.Tensor], spec: ConvSpec, *, relu: bool = False,
                        mask_sigmoid: bool = False, backend: str = "auto",
                        outs: Optional[Sequence[torch.Tensor]] = None, hint: int = 0,
                        out_f32: bool = False, out_planar: bool = False) -> List[torch.Tensor]:
    """One launch over several feature maps that share one weight (e.g. the FPN levels of the
    shared prediction head, reference STMask.py:91-92 / prediction_head_FC.py:166-167).

    offsets[i] is None for every i  =>  plain convolution through the same kernel.
    out_f32: the outputs are float32 whatever the activations' dtype (the fp32 accumulators are stored as they
    are) — what the offset / mask-logit predictor of a DCN wants: sampling positions must not be rounded to bf16.
    out_planar: the outputs are contiguous [B, Cout, Ho, Wo] ("NCHW") tensors instead of channels-last ones (tcgen05
    backend only; silently channels-last when the call runs on the CUDA-core kernel) — the layout in which a sampling
    kernel's per-tap offset loads of neighbouring pixels coalesce.
    """
    n = len(xs)
    if n == 0:
        return []
    _no_grad_inputs(*xs, *[o for o in offsets if o is not None], *([m for m in masks if m is not None] if masks else []))
    if n > L.DCN_MAX_PROBLEMS:
        raise ValueError(f"at most {L.DCN_MAX_PROBLEMS} feature maps per launch, got {n}")
    if len(offsets) != n or (masks is not None and len(masks) != n):
        raise ValueError("xs / offsets / masks length mismatch")
    zero_offset = all(o is None for o in offsets)
    if not zero_offset and any(o is None for o in offsets):
        raise ValueError("either every problem has an offset tensor or none has")
    _require_cuda(w_packed, "packed weight")
    dev = w_packed.device
    kh, kw = spec.kernel
    probs = (L.StmDcnProblem * n)()
    keep = []
    ys: List[torch.Tensor] = []
    xdt = None
    odt = L.STM_F32
    have_mask = None
    for i in range(n):
        x = xs[i]
        _require_cuda(x, "x")
        if x.dim() != 4:
            raise ValueError(f"x must be 4-D (N, C, H, W), got {tuple(x.shape)}")
        if x.device != dev:
            raise ValueError("x and weight are on different devices")
        if x.shape[1] != spec.in_c:
            raise ValueError(f"x has {x.shape[1]} channels, conv expects {spec.in_c}")
        d = _dt(x, "x")
        if xdt is None:
            xdt = d
            if w_packed.dtype != x.dtype:
                raise TypeError(f"packed weight dtype {w_packed.dtype} != x dtype {x.dtype}")
        elif d != xdt:
            raise TypeError("all feature maps of one launch must share a dtype")
        b, _, h, w = x.shape
        ho, wo = spec.out_hw(h, w)
        if ho <= 0 or wo <= 0:
            raise ValueError(f"convolution output size would be {ho}x{wo}")
        xn = to_nhwc(x)
        off = offsets[i]
        msk = masks[i] if masks is not None else None
        if have_mask is None:
            have_mask = msk is not None
        elif have_mask != (msk is not None):
            raise ValueError("either every problem has a mask or none has")
        _check_offset(off, spec.deform_groups * 2 * kh * kw, b, ho, wo, "offset")
        _check_offset(msk, spec.deform_groups * kh * kw, b, ho, wo, "mask")
        first = off if off is not None else msk
        if first is not None:
            d_off = _dt(first, "offset")
            if not keep:
                odt = d_off
            elif d_off != odt:
                raise TypeError("all offsets / masks of one launch must share a dtype")
            if msk is not None and msk.dtype != first.dtype:
                msk = msk.to(first.dtype)
        if outs is not None:
            y = outs[i]
            if tuple(y.shape) != (b, spec.out_c, ho, wo) or y.dtype != (torch.float32 if out_f32 else x.dtype) or \
                    not (spec.out_c == 1 or (y.stride(1) == ho * y.stride(2) if out_planar else y.stride(1) == 1)):
                raise ValueError("preallocated output must be a channels-last (or, with out_planar, plane-major) tensor of the right shape/dtype")
        elif out_planar:
            y = torch.empty((b, spec.out_c, ho, wo), dtype=torch.float32 if out_f32 else x.dtype, device=dev)
        else:
            y = torch.empty((b, spec.out_c, ho, wo), dtype=torch.float32 if out_f32 else x.dtype, device=dev,
                            memory_format=torch.channels_last)
        keep += [xn, off, msk, y]
        ys.append(y)
        p = probs[i]
        p.batch, p.in_h, p.in_w, p.out_h, p.out_w = b, h, w, ho, wo
        p.x = xn.data_ptr()
        p.x_stride_n, p.x_stride_h, p.x_stride_w = xn.stride(0), xn.stride(2), xn.stride(3)
        if off is not None:
            p.offset = off.data_ptr()
            p.off_stride_n, p.off_stride_c, p.off_stride_h, p.off_stride_w = off.stride()
        if msk is not None:
            p.mask = msk.data_ptr()
            p.mask_stride_n, p.mask_stride_c, p.mask_stride_h, p.mask_stride_w = msk.stride()
        p.y = y.data_ptr()
        p.y_stride_n, p.y_stride_h, p.y_stride_w = y.stride(0), y.stride(2), y.stride(3)
    flags = (L.DCN_RELU if relu else 0) | (L.DCN_MASK_SIGMOID if mask_sigmoid else 0) | (L.DCN_ZERO_OFFSET if zero_offset else 0)
    flags |= int(hint) & (L.DCN_HINT_ROWS128 | L.DCN_HINT_ROWS256 | L.DCN_HINT_NO_PAIR | L.DCN_HINT_DEEP_PIPE | L.DCN_HINT_TWO_CTAS | L.DCN_HINT_GATHER | L.DCN_HINT_TAP_MAJOR | L.DCN_HINT_CHUNK_MAJOR | L.DCN_HINT_NO_FUSE | L.DCN_HINT_RASTER | (0xff << 20))
    if out_f32:
        flags |= L.DCN_OUT_F32
    if out_planar:
        flags |= L.DCN_OUT_PLANAR
    conv = spec.c_struct(xdt, odt, flags, _BACKENDS[backend])
    if out_planar and L.lib().stm_deform_conv2d_backend(C.byref(conv), probs, n) != L.BACKEND_TCGEN05:
        # the CUDA-core kernel writes channels-last only: same values, the other memory format
        if outs is not None:
            raise ValueError("out_planar outputs were preallocated but this call runs on the CUDA-core kernel")
        flags &= ~L.DCN_OUT_PLANAR
        conv = spec.c_struct(xdt, odt, flags, _BACKENDS[backend])
        for i in range(n):
            y = torch.empty(tuple(ys[i].shape), dtype=ys[i].dtype, device=dev, memory_format=torch.channels_last)
            ys[i] = y
            keep.append(y)
            probs[i].y = y.data_ptr()
            probs[i].y_stride_n, probs[i].y_stride_h, probs[i].y_stride_w = y.stride(0), y.stride(2), y.stride(3)
    if bias_f32 is not None:
        _require_cuda(bias_f32, "bias")
        if bias_f32.dtype != torch.float32 or tuple(bias_f32.shape) != (spec.out_c,) or not bias_f32.is_contiguous():
            raise ValueError("bias must be a contiguous float32 [out_channels] tensor")
    lib = L.lib()
    with torch.cuda.device(dev):
        ws_bytes = lib.stm_deform_conv2d_workspace(C.byref(conv), probs, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
        rc = lib.stm_deform_conv2d_fwd(C.byref(conv), probs, n, w_packed.data_ptr(),
                                       bias_f32.data_ptr() if bias_f32 is not None else None,
                                       ws.data_ptr() if ws is not None else None, ws_bytes,
                                       torch.cuda.current_stream(dev).cuda_stream)
    L.check(rc, "stm_deform_conv2d_fwd")
    return ys


def deform_conv2d_fcb_multi(xs: Sequence[torch.Tensor], deltas: Sequence[torch.Tensor], w_packed: torch.Tensor, spec: ConvSpec,
                            fcb_weight: Optional[torch.Tensor], *, relu: bool = True,
                            outs: Optional[Sequence[torch.Tensor]] = None, hint: int = 0) -> List[torch.Tensor]:
    """FeatureAlign's deformable conv over several FPN levels in ONE launch with the offsets derived INSIDE the kernel
    from the regressed box deltas (`deltas[i]`: [B, 4, H, W] = (t_x, t_y, t_w, t_h), any strides, fp32 or bf16):
    FCB(ada) when `fcb_weight` (the 1x1 conv_offset weight [dg*2*kh*kw, 4, 1, 1]) is given, FCB(ali) (closed form)
    when it is None.  tcgen05 backend only — raises StmError (status -2) for shapes that need the CUDA-core kernel."""
    n = len(xs)
    if n == 0:
        return []
    if n > L.DCN_MAX_PROBLEMS or len(deltas) != n:
        raise ValueError("xs / deltas length mismatch or too many feature maps")
    _no_grad_inputs(*xs, *deltas)
    dev = w_packed.device
    probs = (L.StmDcnProblem * n)()
    keep, ys = [], []
    odt = None
    for i in range(n):
        x, dl = xs[i], deltas[i]
        _require_cuda(x, "x")
        _require_cuda(dl, "box deltas")
        if x.dim() != 4 or x.shape[1] != spec.in_c:
            raise ValueError(f"x must be [B, {spec.in_c}, H, W], got {tuple(x.shape)}")
        b, _, h, w = x.shape
        ho, wo = spec.out_hw(h, w)
        if tuple(dl.shape) != (b, 4, ho, wo):
            raise ValueError(f"box deltas shape {tuple(dl.shape)} != {(b, 4, ho, wo)}")
        d_off = _dt(dl, "box deltas")
        if odt is None:
            odt = d_off
        elif odt != d_off:
            raise TypeError("all box-delta tensors of one launch must share a dtype")
        xn = to_nhwc(x)
        if outs is not None:
            y = outs[i]
            if tuple(y.shape) != (b, spec.out_c, ho, wo) or y.dtype != x.dtype or y.stride(1) != 1:
                raise ValueError("preallocated output must be a channels-last tensor of the right shape/dtype")
        else:
            y = torch.empty((b, spec.out_c, ho, wo), dtype=x.dtype, device=dev, memory_format=torch.channels_last)
        keep += [xn, dl, y]
        ys.append(y)
        p = probs[i]
        p.batch, p.in_h, p.in_w, p.out_h, p.out_w = b, h, w, ho, wo
        p.x = xn.data_ptr()
        p.x_stride_n, p.x_stride_h, p.x_stride_w = xn.stride(0), xn.stride(2), xn.stride(3)
        p.offset = dl.data_ptr()
        p.off_stride_n, p.off_stride_c, p.off_stride_h, p.off_stride_w = dl.stride()
        p.y = y.data_ptr()
        p.y_stride_n, p.y_stride_h, p.y_stride_w = y.stride(0), y.stride(2), y.stride(3)
    fw = None
    if fcb_weight is not None:
        fw = fcb_weight.detach().reshape(-1, 4).float().contiguous()
        if fw.shape[0] != spec.deform_groups * 2 * spec.kernel[0] * spec.kernel[1]:
            raise ValueError("conv_offset weight does not match kernel_size / deform_groups")
    flags = (L.DCN_RELU if relu else 0) | (L.DCN_FCB_ADA if fw is not None else L.DCN_FCB_ALI)
    flags |= int(hint) & (L.DCN_HINT_ROWS128 | L.DCN_HINT_ROWS256 | L.DCN_HINT_NO_PAIR | L.DCN_HINT_TWO_CTAS | L.DCN_HINT_TAP_MAJOR | L.DCN_HINT_CHUNK_MAJOR | L.DCN_HINT_RASTER)
    conv = spec.c_struct(_dt(xs[0], "x"), odt, flags, L.BACKEND_AUTO)
    with torch.cuda.device(dev):
        rc = L.lib().stm_deform_conv2d_fcb_fwd(C.byref(conv), probs, n, w_packed.data_ptr(), None,
                                               fw.data_ptr() if fw is not None else None, None, 0,
                                               torch.cuda.current_stream(dev).cuda_stream)
    L.check(rc, "stm_deform_conv2d_fcb_fwd")
    return ys


def deform_conv2d_backend(x_shape: Sequence[int], spec: ConvSpec, dtype: torch.dtype, backend: str = "auto") -> str:
    """Which kernel family a call would use ('simt' / 'tcgen05') — shape-only query."""
    b, _, h, w = x_shape
    ho, wo = spec.out_hw(h, w)
    prob = (L.StmDcnProblem * 1)()
    p = prob[0]
    p.batch, p.in_h, p.in_w, p.out_h, p.out_w = b, h, w, ho, wo
    p.x = p.y = p.offset = 256           # dummy non-null, 256-byte aligned addresses; never dereferenced here
    p.x_stride_w, p.x_stride_h, p.x_stride_n = spec.in_c, spec.in_c * w, spec.in_c * w * h
    p.y_stride_w, p.y_stride_h, p.y_stride_n = spec.out_c, spec.out_c * wo, spec.out_c * wo * ho
    dt = L.STM_F32 if dtype == torch.float32 else L.STM_BF16
    conv = spec.c_struct(dt, L.STM_F32, 0, _BACKENDS[backend])
    rc = L.lib().stm_deform_conv2d_backend(C.byref(conv), prob, 1)
    if rc < 0:
        L.check(rc, "stm_deform_conv2d_backend")
    return L.BACKEND_NAMES[rc]


def deform_conv2d_variant(x_shapes: Sequence[Sequence[int]], spec: ConvSpec, dtype: torch.dtype, backend: str = "auto",
                          hint: int = 0, device=None, zero_offset: bool = False, fcb: bool = False) -> str:
    """The kernel instantiation a launch over feature maps of these shapes would run on `device`
    (e.g. 'tcgen05 rows=256 n=256 producer_warps=16 stages=2 pair=1 ...') — shape-only query."""
    n = len(x_shapes)
    prob = (L.StmDcnProblem * n)()
    for p, (b, _, h, w) in zip(prob, x_shapes):
        ho, wo = spec.out_hw(h, w)
        p.batch, p.in_h, p.in_w, p.out_h, p.out_w = b, h, w, ho, wo
        p.x = p.y = p.offset = 256
        p.x_stride_w, p.x_stride_h, p.x_stride_n = spec.in_c, spec.in_c * w, spec.in_c * w * h
        p.y_stride_w, p.y_stride_h, p.y_stride_n = spec.out_c, spec.out_c * wo, spec.out_c * wo * ho
    dt = L.STM_F32 if dtype == torch.float32 else L.STM_BF16
    conv = spec.c_struct(dt, L.STM_F32, int(hint) | (L.DCN_ZERO_OFFSET if zero_offset else 0) | (L.DCN_FCB_ADA if fcb else 0),
                         _BACKENDS[backend])
    buf = C.create_string_buffer(256)
    import contextlib
    ctx = torch.cuda.device(device if device is not None else torch.cuda.current_device()) if torch.cuda.is_available() \
        else contextlib.nullcontext()      # without a GPU the plan assumes a B200's 148 SMs
    with ctx:
        L.check(L.lib().stm_deform_conv2d_variant(C.byref(conv), prob, n, buf, 256), "stm_deform_conv2d_variant")
    return buf.value.decode()


_functional_cache = PackedWeightCache(max_entries=64)


def deform_conv2d(x: torch.Tensor, offset: Optional[torch.Tensor], weight: torch.Tensor,
                  bias: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, stride: IntPair = 1,
                  padding: IntPair = 0, dilation: IntPair = 1, groups: int = 1, deform_groups: int = 1, *,
                  relu: bool = False, mask_sigmoid: bool = False, backend: str = "auto",
                  cache: Optional[PackedWeightCache] = None, hint: int = 0) -> torch.Tensor:
    """DCNv1 (mask None) / DCNv2 forward with an OIHW weight.  The packed copy of the weight is cached (the
    functional drop-ins `dcn_v2_conv`, `mmcv.ops.deform_conv2d`, `modulated_deform_conv2d` therefore pack once
    per weight tensor, not once per call); entries are tied to the weight OBJECT by weak reference, so they
    can neither alias a new tensor at a recycled address nor outlive the weight."""
    _require_cuda(x, "x")
    _require_cuda(weight, "weight")
    _no_grad_inputs(weight, bias)
    spec = ConvSpec(x.shape[1] if x.dim() == 4 else -1, weight.shape[0], weight.shape[2:], stride, padding, dilation,
                    groups, deform_groups)
    cache = cache if cache is not None else _functional_cache
    wp = cache.weight(weight, spec, x.dtype)
    bf = cache.bias(bias)
    return deform_conv2d_multi([x], [offset], [mask] if mask is not None else None, wp, bf, spec, relu=relu,
                               mask_sigmoid=mask_sigmoid, backend=backend, hint=hint)[0]


class PlainConv:
    """A regular convolution (+ bias, optional ReLU) run through the deformable-conv kernels in their zero-offset
    mode: tcgen05 implicit GEMM with copy-only producers for bf16 NHWC activations whose channel count is a multiple
    of 64, the CUDA-core kernel otherwise.  Used for the offset / mask-logit predictor of `DCN` (fp32 output) and for
    the TemporalNet convs.  Holds the packed weight (output channels zero-padded to a multiple of 16) and the fp32
    bias of ONE nn.Conv2d-like parameter pair, re-packed when the parameters change."""

    def __init__(self):
        self._key = None
        self._packed = None

    def _refresh(self, weight: torch.Tensor, bias: Optional[torch.Tensor], dtype: torch.dtype, in_pad: Optional[Tuple[int, int]]):
        key = (id(weight), weight._version, weight.data_ptr(), weight.dtype, dtype, in_pad,
               None if bias is None else (id(bias), bias._version, bias.data_ptr()))
        if key != self._key:
            co = weight.shape[0]
            cp = (co + 15) // 16 * 16
            w = weight.detach()
            if in_pad is not None:            # insert zero input channels at [at, at + n) (padded concat layout)
                at, n = in_pad
                w = torch.cat([w[:, :at], w.new_zeros((co, n) + tuple(w.shape[2:])), w[:, at:]], dim=1)
            if cp != co:
                w = torch.cat([w, w.new_zeros((cp - co,) + tuple(w.shape[1:]))], dim=0)
            spec = ConvSpec(w.shape[1], cp, w.shape[2:])
            b = None
            if bias is not None:
                b = torch.zeros(cp, dtype=torch.float32, device=weight.device)
                b[:co] = bias.detach().float()
            self._packed = (pack_weight(w.contiguous(), spec, dtype), b, cp)
            self._key = key
        return self._packed

    def __call__(self, xs: Sequence[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor], stride: IntPair = 1,
                 padding: IntPair = 0, dilation: IntPair = 1, *, relu: bool = False, out_f32: bool = False,
                 in_pad: Optional[Tuple[int, int]] = None, backend: str = "auto", hint: int = 0,
                 out_planar: bool = False) -> List[torch.Tensor]:
        """-> list of [B, Cout_padded, Ho, Wo] channels-last tensors (slice [:, :Cout] is the convolution); with
        `out_planar` contiguous plane-major ("NCHW") tensors when the call runs on the tcgen05 kernels."""
        _no_grad_inputs(weight, bias, *xs)
        wp, b, cp = self._refresh(weight, bias, xs[0].dtype, in_pad)
        spec = ConvSpec(xs[0].shape[1], cp, weight.shape[2:], stride, padding, dilation)
        return deform_conv2d_multi(list(xs), [None] * len(xs), None, wp, b, spec, relu=relu, backend=backend, hint=hint,
                                   out_f32=out_f32, out_planar=out_planar)


def fcb_ali_offsets(shape: torch.Tensor, kernel_size: IntPair, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Closed-form FCB(ali) offsets from box deltas (reference Featurealign.py:46-69)."""
    _require_cuda(shape, "shape")
    if shape.dim() != 4 or shape.shape[1] != 4:
        raise ValueError(f"shape must be [B, 4, H, W], got {tuple(shape.shape)}")
    kh, kw = _pair(kernel_size)
    b, _, h, w = shape.shape
    off = torch.empty((b, 2 * kh * kw, h, w), dtype=dtype or shape.dtype, device=shape.device)
    if off.numel() == 0:
        return off
    ss = (C.c_int64 * 4)(*shape.stride())
    os_ = (C.c_int64 * 4)(*off.stride())
    with torch.cuda.device(shape.device):
        L.check(L.lib().stm_fcb_ali_offsets(shape.data_ptr(), ss, _dt(shape, "shape"), off.data_ptr(), os_, _dt(off, "offset"),
                                            b, h, w, kh, kw, _stream(shape)), "stm_fcb_ali_offsets")
    return off


def fcb_ada_offsets(shape: torch.Tensor, weight: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """FCB(ada) offsets: the bias-free 1x1 `conv_offset` applied to the box deltas
    (reference Featurealign.py:20-25,44).  weight: [OC, 4, 1, 1]."""
    _require_cuda(shape, "shape")
    _require_cuda(weight, "conv_offset.weight")
    if shape.dim() != 4 or shape.shape[1] != 4:
        raise ValueError(f"shape must be [B, 4, H, W], got {tuple(shape.shape)}")
    if weight.dim() != 4 or tuple(weight.shape[1:]) != (4, 1, 1):
        raise ValueError(f"conv_offset.weight must be [OC, 4, 1, 1], got {tuple(weight.shape)}")
    b, _, h, w = shape.shape
    oc = weight.shape[0]
    w2 = weight.detach().reshape(oc, 4).float().contiguous()
    off = torch.empty((b, oc, h, w), dtype=dtype or shape.dtype, device=shape.device)
    if off.numel() == 0:
        return off
    ss = (C.c_int64 * 4)(*shape.stride())
    os_ = (C.c_int64 * 4)(*off.stride())
    with torch.cuda.device(shape.device):
        L.check(L.lib().stm_fcb_ada_offsets(shape.data_ptr(), ss, _dt(shape, "shape"), w2.data_ptr(), off.data_ptr(), os_,
                                            _dt(off, "offset"), b, h, w, oc, _stream(shape)), "stm_fcb_ada_offsets")
    return off


# --------------------------------------------------------------------------------------------
# correlation
# --------------------------------------------------------------------------------------------
def correlation(x1: torch.Tensor, x2: torch.Tensor, patch_size: int = 11, dilation_patch: int = 1, *,
                scale: float = 1.0, leaky_slope: Optional[float] = None, relu: bool = False,
                feats: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, out_dtype: Optional[torch.dtype] = None,
                channels_last: bool = False, backend: str = "auto", feat_channel_offset: Optional[int] = None) -> torch.Tensor:
    """Cost volume [B, P*P, H, W] (+ 2*feat_c concat channels when `feats` is given).

    `feat_channel_offset` (with `feats`): channel at which the first feature map starts (default P*P, the
    reference's concat).  A larger value pads the correlation block with zero channels; 128 for P = 11 makes
    every block of a channels-last bf16 pixel row 16-byte aligned, which is what the fast epilogue needs.

    out[b, ph*P+pw, y, x] = post(scale * <x1[b,:,y,x], x2[b,:,y+(ph-r)d, x+(pw-r)d]>), zero outside x2.
    `channels_last=True` returns NHWC memory (what the RoIAlign/TemporalNet convs want);
    the default is contiguous NCHW so that the reference's `.view(b, ph*pw, h, w)` works.
    """
    _require_cuda(x1, "input1")
    _require_cuda(x2, "input2")
    if x1.dim() != 4 or x1.shape != x2.shape:
        raise ValueError(f"input1/input2 must be 4-D of equal shape, got {tuple(x1.shape)} and {tuple(x2.shape)}")
    if x1.dtype != x2.dtype or x1.device != x2.device:
        raise TypeError("input1/input2 dtype or device mismatch")
    P, d = int(patch_size), int(dilation_patch)
    if P < 1 or P % 2 == 0:
        raise ValueError(f"patch_size must be odd and positive, got {P}")
    if d < 1:
        raise ValueError("dilation_patch must be >= 1")
    b, c, h, w = x1.shape
    a, bb = to_nhwc(x1), to_nhwc(x2)
    flags = 0
    fa = fb = None
    fc = 0
    fdt = L.STM_BF16
    if leaky_slope is not None:
        flags |= L.CORR_LEAKY_RELU
    if relu:
        flags |= L.CORR_RELU
    if feats is not None:
        fa, fb = feats
        _require_cuda(fa, "feat_a")
        _require_cuda(fb, "feat_b")
        if fa.shape != fb.shape or fa.dim() != 4 or fa.shape[0] != b or tuple(fa.shape[2:]) != (h, w) or fa.dtype != fb.dtype:
            raise ValueError("feats must be two [B, Cf, H, W] tensors matching the inputs' batch and size")
        fa, fb = to_nhwc(fa), to_nhwc(fb)
        fc = fa.shape[1]
        fdt = _dt(fa, "feats")
        flags |= L.CORR_COPY_FEATS
    odt = out_dtype or x1.dtype
    foff = P * P
    if feat_channel_offset is not None:
        if feats is None or int(feat_channel_offset) < P * P:
            raise ValueError("feat_channel_offset needs feats and must be >= patch_size**2")
        foff = int(feat_channel_offset)
    ch = foff + 2 * fc
    out = torch.empty((b, ch, h, w), dtype=odt, device=x1.device,
                      memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    if out.numel() == 0:
        return out
    desc = L.StmCorrDesc()
    desc.batch, desc.h, desc.w, desc.c = b, h, w, c
    desc.patch, desc.dilation_patch = P, d
    desc.dtype, desc.out_dtype = _dt(a, "input1"), _dt(out, "out")
    desc.flags, desc.backend = flags, _BACKENDS[backend]
    desc.scale, desc.leaky_slope = float(scale), float(leaky_slope or 0.0)
    desc.x1_stride_n, desc.x1_stride_h, desc.x1_stride_w = a.stride(0), a.stride(2), a.stride(3)
    desc.x2_stride_n, desc.x2_stride_h, desc.x2_stride_w = bb.stride(0), bb.stride(2), bb.stride(3)
    desc.out_stride_n, desc.out_stride_c, desc.out_stride_h, desc.out_stride_w = out.stride()
    desc.feat_c, desc.feat_dtype = fc, fdt
    desc.feat_c_offset = foff if feats is not None else 0
    if fa is not None:
        desc.feat_a_stride_n, desc.feat_a_stride_h, desc.feat_a_stride_w = fa.stride(0), fa.stride(2), fa.stride(3)
        desc.feat_b_stride_n, desc.feat_b_stride_h, desc.feat_b_stride_w = fb.stride(0), fb.stride(2), fb.stride(3)
    with torch.cuda.device(x1.device):
        rc = L.lib().stm_correlation_fwd(C.byref(desc), a.data_ptr(), bb.data_ptr(),
                                         fa.data_ptr() if fa is not None else None,
                                         fb.data_ptr() if fb is not None else None, out.data_ptr(), _stream(x1))
    L.check(rc, "stm_correlation_fwd")
    return out


def correlation_multi(x1s: Sequence[torch.Tensor], x2s: Sequence[torch.Tensor], patch_size: int = 11, dilation_patch: int = 1, *,
                      scale: float = 1.0, leaky_slope: Optional[float] = None, relu: bool = False,
                      padded: bool = True) -> List[torch.Tensor]:
    """Cost volumes of several feature maps (the FPN levels P3..P7) in ONE launch: out[i] is the channels-last bf16
    cost volume of (x1s[i], x2s[i]).  `padded` (default): rows of padded_channels = ceil(P*P / 8) * 8 channels (128 for
    P = 11; channels [P*P, 128) are zeros) so that every pixel row is 256 bytes and is written with 16-byte stores —
    `out[i][:, :P*P]` is the reference's `[B, P*P, H, W]` view; `padded=False`: exactly P*P channels (element-wise stores).
    bf16 / tcgen05 only."""
    n = len(x1s)
    if n == 0:
        return []
    if n != len(x2s) or n > 8:
        raise ValueError("x1s / x2s must have the same length (<= 8)")
    P, d = int(patch_size), int(dilation_patch)
    if P < 1 or P % 2 == 0 or d < 1:
        raise ValueError("patch_size must be odd and positive, dilation_patch >= 1")
    flags = (L.CORR_LEAKY_RELU if leaky_slope is not None else 0) | (L.CORR_RELU if relu else 0)
    ch = (P * P + 7) // 8 * 8 if padded else P * P
    descs = (L.StmCorrDesc * n)()
    p1, p2, po = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
    keep, outs = [], []
    for i, (x1, x2) in enumerate(zip(x1s, x2s)):
        _require_cuda(x1, "input1")
        _require_cuda(x2, "input2")
        if x1.dim() != 4 or x1.shape != x2.shape or x1.dtype != torch.bfloat16 or x2.dtype != torch.bfloat16:
            raise TypeError("correlation_multi takes pairs of equal-shape 4-D bf16 tensors")
        a, b = to_nhwc(x1), to_nhwc(x2)
        bsz, c, h, w = a.shape
        out = torch.empty((bsz, ch, h, w), dtype=torch.bfloat16, device=a.device, memory_format=torch.channels_last)
        keep += [a, b]
        outs.append(out)
        q = descs[i]
        q.batch, q.h, q.w, q.c = bsz, h, w, c
        q.patch, q.dilation_patch = P, d
        q.dtype = q.out_dtype = L.STM_BF16
        q.flags, q.backend = flags, L.BACKEND_TCGEN05
        q.scale, q.leaky_slope = float(scale), float(leaky_slope or 0.0)
        q.x1_stride_n, q.x1_stride_h, q.x1_stride_w = a.stride(0), a.stride(2), a.stride(3)
        q.x2_stride_n, q.x2_stride_h, q.x2_stride_w = b.stride(0), b.stride(2), b.stride(3)
        q.out_stride_n, q.out_stride_c, q.out_stride_h, q.out_stride_w = out.stride()
        q.feat_c_offset = ch if padded else 0
        p1[i], p2[i], po[i] = a.data_ptr(), b.data_ptr(), out.data_ptr()
    dev = outs[0].device
    with torch.cuda.device(dev):
        rc = L.lib().stm_correlation_multi_fwd(descs, p1, p2, po, n, torch.cuda.current_stream(dev).cuda_stream)
    L.check(rc, "stm_correlation_multi_fwd")
    return outs


_checked_indices = {}


def _check_pair_indices(ref_index: torch.Tensor, next_index: torch.Tensor, frames: int, halo_frames: int) -> None:
    """Range-check the pair index arrays ONCE per (tensor, version): 0 <= next < F and 0 <= ref < F + halo frames
    (a ref >= F addresses the halo, so a halo must be present).  The check reads the indices back to the host,
    which is why the verdict is remembered for the (persistent, cached) index tensors sharding.py hands out."""
    if ref_index.numel() == 0:
        return
    import weakref
    key = (id(ref_index), id(next_index))
    hit = _checked_indices.get(key)
    sig = (ref_index._version, next_index._version, frames, halo_frames)
    if hit is not None and hit[0]() is ref_index and hit[1]() is next_index and hit[2] == sig:
        return
    lo = int(torch.minimum(ref_index.min(), next_index.min()))
    rmax, nmax = int(ref_index.max()), int(next_index.max())
    if lo < 0 or nmax >= frames or rmax >= frames + halo_frames:
        raise ValueError(f"pair indices out of range: ref in [{lo}, {rmax}], next max {nmax}, {frames} frames + {halo_frames} "
                         f"halo frames (a ref index >= {frames} needs a halo tensor)")
    if len(_checked_indices) > 64:
        _checked_indices.clear()
    _checked_indices[key] = (weakref.ref(ref_index), weakref.ref(next_index), sig)


def correlation_pairs(x: torch.Tensor, ref_index: torch.Tensor, next_index: torch.Tensor, patch_size: int = 11,
                      dilation_patch: int = 1, *, scale: float = 1.0, leaky_slope: Optional[float] = None, relu: bool = False,
                      feats: Optional[torch.Tensor] = None, halo: Optional[torch.Tensor] = None,
                      feats_halo: Optional[torch.Tensor] = None, feat_channel_offset: Optional[int] = None) -> torch.Tensor:
    """Temporal-fusion cost volume straight out of a FRAME BATCH: pair i correlates frame ref_index[i] with frame
    next_index[i] of `x` [F, C, H, W] (and, with `feats` [F, Ct, H, W], concatenates feats[ref_index[i]],
    feats[next_index[i]] behind it) — no gathered copies of the (t-1, t) pairs.  A ref_index value v >= F addresses
    frame v - F of `halo` / `feats_halo`: the one-frame halos received from the neighbour rank (sharding.py).
    Indices are int32 device tensors.  bf16 / tcgen05 only; channels-last output [n_pairs, Cout, H, W]."""
    _require_cuda(x, "x")
    if x.dim() != 4:
        raise ValueError("x must be [F, C, H, W]")
    if ref_index.dtype != torch.int32 or next_index.dtype != torch.int32 or ref_index.shape != next_index.shape or ref_index.dim() != 1:
        raise TypeError("ref_index / next_index must be 1-D int32 tensors of equal length")
    if not ref_index.is_cuda or not next_index.is_cuda:
        raise RuntimeError("ref_index / next_index must be CUDA tensors")
    if x.dtype != torch.bfloat16:
        raise TypeError("correlation_pairs runs on the tcgen05 backend: bf16 only")
    P, d = int(patch_size), int(dilation_patch)
    if P < 1 or P % 2 == 0 or d < 1:
        raise ValueError("patch_size must be odd and positive, dilation_patch >= 1")
    f, c, h, w = x.shape
    n = ref_index.numel()
    n_halo = int(halo.shape[0]) if halo is not None else 0
    _check_pair_indices(ref_index, next_index, f, n_halo)
    a = to_nhwc(x)
    flags = 0
    if leaky_slope is not None:
        flags |= L.CORR_LEAKY_RELU
    if relu:
        flags |= L.CORR_RELU
    fc = 0
    ft = None
    if feats is not None:
        _require_cuda(feats, "feats")
        if feats.dim() != 4 or feats.shape[0] != f or tuple(feats.shape[2:]) != (h, w) or feats.dtype != x.dtype:
            raise ValueError("feats must be [F, Ct, H, W] matching x")
        ft = to_nhwc(feats)
        fc = ft.shape[1]
        flags |= L.CORR_COPY_FEATS
    ha = hf = None
    if halo is not None and halo.shape[0] > 0:
        if tuple(halo.shape[1:]) != (c, h, w) or halo.dtype != x.dtype:
            raise ValueError("halo must be [Fh, C, H, W] matching x")
        ha = to_nhwc(halo)
        if feats is not None:
            if feats_halo is None or feats_halo.shape[0] != halo.shape[0] or tuple(feats_halo.shape[1:]) != (fc, h, w):
                raise ValueError("feats_halo must be [Fh, Ct, H, W]")
            hf = to_nhwc(feats_halo.to(x.dtype))
    foff = P * P
    if feat_channel_offset is not None:
        if feats is None or int(feat_channel_offset) < P * P:
            raise ValueError("feat_channel_offset needs feats and must be >= patch_size**2")
        foff = int(feat_channel_offset)
    out = torch.empty((n, foff + 2 * fc if feats is not None else P * P, h, w), dtype=x.dtype, device=x.device,
                      memory_format=torch.channels_last)
    if out.numel() == 0:
        return out
    desc = L.StmCorrDesc()
    desc.batch, desc.h, desc.w, desc.c = n, h, w, c
    desc.patch, desc.dilation_patch = P, d
    desc.dtype = desc.out_dtype = L.STM_BF16
    desc.flags, desc.backend = flags, L.BACKEND_TCGEN05
    desc.scale, desc.leaky_slope = float(scale), float(leaky_slope or 0.0)
    desc.x1_stride_n, desc.x1_stride_h, desc.x1_stride_w = a.stride(0), a.stride(2), a.stride(3)
    desc.x2_stride_n, desc.x2_stride_h, desc.x2_stride_w = a.stride(0), a.stride(2), a.stride(3)
    desc.out_stride_n, desc.out_stride_c, desc.out_stride_h, desc.out_stride_w = out.stride()
    desc.feat_c, desc.feat_dtype = fc, L.STM_BF16
    desc.feat_c_offset = foff if feats is not None else 0
    if ft is not None:
        desc.feat_a_stride_n, desc.feat_a_stride_h, desc.feat_a_stride_w = ft.stride(0), ft.stride(2), ft.stride(3)
        desc.feat_b_stride_n, desc.feat_b_stride_h, desc.feat_b_stride_w = ft.stride(0), ft.stride(2), ft.stride(3)
    desc.x1_index, desc.x2_index = ref_index.data_ptr(), next_index.data_ptr()
    desc.x1_frames = desc.x2_frames = f
    if ha is not None:
        desc.alt_frames = ha.shape[0]
        desc.x1_alt = ha.data_ptr()
        desc.x1_alt_stride_n, desc.x1_alt_stride_h, desc.x1_alt_stride_w = ha.stride(0), ha.stride(2), ha.stride(3)
        if hf is not None:
            desc.feat_a_alt = hf.data_ptr()
            desc.feat_a_alt_stride_n, desc.feat_a_alt_stride_h, desc.feat_a_alt_stride_w = hf.stride(0), hf.stride(2), hf.stride(3)
    with torch.cuda.device(x.device):
        rc = L.lib().stm_correlation_fwd(C.byref(desc), a.data_ptr(), a.data_ptr(), ft.data_ptr() if ft is not None else None,
                                         ft.data_ptr() if ft is not None else None, out.data_ptr(), _stream(x))
    L.check(rc, "stm_correlation_fwd")
    return out


def correlation_backend(shape: Sequence[int], dtype: torch.dtype, patch_size: int = 11, dilation_patch: int = 1,
                        backend: str = "auto") -> str:
    b, c, h, w = shape
    desc = L.StmCorrDesc()
    desc.batch, desc.h, desc.w, desc.c = b, h, w, c
    desc.patch, desc.dilation_patch = patch_size, dilation_patch
    desc.dtype = desc.out_dtype = L.STM_F32 if dtype == torch.float32 else L.STM_BF16
    desc.backend = _BACKENDS[backend]
    desc.x1_stride_w = desc.x2_stride_w = c
    desc.x1_stride_h = desc.x2_stride_h = c * w
    desc.x1_stride_n = desc.x2_stride_n = c * w * h
    desc.out_stride_w, desc.out_stride_h, desc.out_stride_c = 1, w, h * w
    desc.out_stride_n = h * w * patch_size * patch_size
    rc = L.lib().stm_correlation_backend(C.byref(desc))
    if rc < 0:
        L.check(rc, "stm_correlation_backend")
    return L.BACKEND_NAMES[rc]


# --------------------------------------------------------------------------------------------
# RoIAlign
# --------------------------------------------------------------------------------------------
def roi_align(input: torch.Tensor, rois: torch.Tensor, output_size=7, spatial_scale: float = 1.0, sampling_ratio: int = 0,
              aligned: bool = True, channels_last: bool = True) -> torch.Tensor:
    """mmcv.ops.roi_align (average pooling) on a CUDA feature map: input [B, C, H, W] (NHWC memory is consumed in
    place), rois [n, 5] = (batch index, x1, y1, x2, y2) -> [n, C, ph, pw] (channels-last memory by default: the
    TemporalNet convs that consume it want NHWC; reference track_to_segment_head.py:65-88)."""
    _require_cuda(input, "input")
    _require_cuda(rois, "rois")
    if input.dim() != 4:
        raise ValueError(f"input must be [B, C, H, W], got {tuple(input.shape)}")
    if rois.dim() != 2 or rois.shape[1] != 5:
        raise ValueError(f"rois must be [n, 5] = (batch index, x1, y1, x2, y2), got {tuple(rois.shape)}")
    ph, pw = _pair(output_size)
    if ph < 1 or pw < 1:
        raise ValueError("output_size must be positive")
    x = to_nhwc(input)
    r = rois.to(torch.float32).contiguous()
    b, c, h, w = x.shape
    n = r.shape[0]
    out = torch.empty((n, c, ph, pw), dtype=input.dtype, device=input.device,
                      memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    if out.numel() == 0:
        return out
    d = L.StmRoiAlignDesc()
    d.batch, d.h, d.w, d.c, d.n_rois = b, h, w, c, n
    d.pooled_h, d.pooled_w = ph, pw
    d.sampling_ratio, d.aligned = int(sampling_ratio), int(bool(aligned))
    d.dtype = d.out_dtype = _dt(x, "input")
    d.spatial_scale = float(spatial_scale)
    d.feat_stride_n, d.feat_stride_h, d.feat_stride_w = x.stride(0), x.stride(2), x.stride(3)
    d.out_stride_n, d.out_stride_c, d.out_stride_h, d.out_stride_w = out.stride()
    with torch.cuda.device(input.device):
        rc = L.lib().stm_roi_align_fwd(C.byref(d), x.data_ptr(), r.data_ptr(), out.data_ptr(), _stream(input))
    L.check(rc, "stm_roi_align_fwd")
    return out


def detect_fast_nms(conf: torch.Tensor, loc: torch.Tensor, centerness: Optional[torch.Tensor], priors: torch.Tensor, *,
                    conf_thresh: float = 0.05, nms_thresh: float = 0.5, top_k: int = 200):
    """Candidate generation + cross-class fast NMS for a batch of frames in one launch (TF_utils.py:54-82,
    detection_TF.py:85-134).  conf [F, P, C] class probabilities, loc [F, P, 4], centerness [F, P] / [F, P, 1] or None,
    priors [P, 4] / [1, P, 4].  Returns (count [F] int32, index [F, top_k] int32, cls [F, top_k] int32 (1-based),
    score [F, top_k], box [F, top_k, 4]) — device tensors; rows are valid up to count[f]; nothing is synchronised."""
    _require_cuda(conf, "conf")
    if conf.dim() != 3 or loc.shape != conf.shape[:2] + (4,):
        raise ValueError("conf must be [F, P, C] and loc [F, P, 4]")
    f, p, c = conf.shape
    pri = priors.reshape(-1, 4)
    if pri.shape[0] != p:
        raise ValueError(f"priors must hold {p} boxes")
    cf, lc, pr = conf.float().contiguous(), loc.float().contiguous(), pri.float().contiguous()
    ct = centerness.reshape(f, p).float().contiguous() if centerness is not None else None
    dev = conf.device
    count = torch.zeros(f, dtype=torch.int32, device=dev)
    index = torch.full((f, top_k), -1, dtype=torch.int32, device=dev)
    cls = torch.zeros((f, top_k), dtype=torch.int32, device=dev)
    score = torch.zeros((f, top_k), dtype=torch.float32, device=dev)
    box = torch.zeros((f, top_k, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.lib().stm_detect_fast_nms_fwd(cf.data_ptr(), lc.data_ptr(), ct.data_ptr() if ct is not None else None, pr.data_ptr(), f, p, c,
                                             int(top_k), float(conf_thresh), float(nms_thresh), count.data_ptr(), index.data_ptr(),
                                             cls.data_ptr(), score.data_ptr(), box.data_ptr(), _stream(conf))
    L.check(rc, "stm_detect_fast_nms_fwd")
    return count, index, cls, score, box


def mask_assembly(proto: torch.Tensor, coeff: torch.Tensor, boxes: torch.Tensor, count: Optional[torch.Tensor] = None):
    """generate_mask + crop for a batch of frames (mask_utils.py:111-128, box_utils.py:341-364): proto [F, h, w, k],
    coeff [F, n, k] (raw; tanh applied inside), boxes [F, n, 4] relative xyxy, count [F] int32 or None.
    Returns (masks [F, n, h, w] float32, mask_bits [F, n, ceil(h*w/32)] int32: masks > 0.5 as bit planes)."""
    _require_cuda(proto, "proto")
    if proto.dim() != 4 or coeff.dim() != 3 or boxes.shape != coeff.shape[:2] + (4,) or coeff.shape[2] != proto.shape[3]:
        raise ValueError("proto [F, h, w, k], coeff [F, n, k], boxes [F, n, 4] expected")
    f, h, w, k = proto.shape
    n = coeff.shape[1]
    words = (h * w + 31) // 32
    pr, cf, bx = proto.float().contiguous(), coeff.float().contiguous(), boxes.float().contiguous()
    masks = torch.zeros((f, n, h, w), dtype=torch.float32, device=proto.device)
    bits = torch.zeros((f, n, words), dtype=torch.int32, device=proto.device)
    with torch.cuda.device(proto.device):
        L.check(L.lib().stm_mask_assembly_fwd(pr.data_ptr(), cf.data_ptr(), bx.data_ptr(), count.data_ptr() if count is not None else None,
                                              masks.data_ptr(), bits.data_ptr(), f, h, w, k, n, _stream(proto)), "stm_mask_assembly_fwd")
    return masks, bits


def mask_iou_bits(bits_a: torch.Tensor, bits_b: torch.Tensor, count_a: Optional[torch.Tensor] = None,
                  count_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mask_iou (box_utils.py:435-447) on bit-plane masks from `mask_assembly`: [F, na, words] x [F, nb, words] -> [F, na, nb]."""
    _require_cuda(bits_a, "bits_a")
    if bits_a.dim() != 3 or bits_b.dim() != 3 or bits_a.shape[0] != bits_b.shape[0] or bits_a.shape[2] != bits_b.shape[2]:
        raise ValueError("bit masks must be [F, n, words] with equal F and words")
    a, b = bits_a.contiguous(), bits_b.contiguous()
    f, na, words = a.shape
    nb = b.shape[1]
    iou = torch.zeros((f, na, nb), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        L.check(L.lib().stm_mask_iou_fwd(a.data_ptr(), b.data_ptr(), count_a.data_ptr() if count_a is not None else None,
                                         count_b.data_ptr() if count_b is not None else None, iou.data_ptr(), f, na, nb, words, _stream(a)),
                "stm_mask_iou_fwd")
    return iou


def pool_fc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """mean over the spatial positions of x [n, C, h, w] (channels-last memory), then y = W * pooled + b in fp32:
    the AvgPool2d(7x7) + Linear tail of TemporalNet (track_to_segment_head.py:17-19,31-35).  weight [out, C], bias [out]."""
    _require_cuda(x, "x")
    _require_cuda(weight, "weight")
    if x.dim() != 4:
        raise ValueError(f"x must be [n, C, h, w], got {tuple(x.shape)}")
    xn = to_nhwc(x)
    n, c, h, w = xn.shape
    if weight.dim() != 2 or weight.shape[1] != c:
        raise ValueError(f"weight must be [out, {c}], got {tuple(weight.shape)}")
    if xn.numel() > 0 and xn.stride(2) != w * xn.stride(3):
        xn = xn.contiguous(memory_format=torch.channels_last)
    wf = weight.detach().float().contiguous()
    bf = bias.detach().float().contiguous() if bias is not None else None
    y = torch.empty((n, wf.shape[0]), dtype=torch.float32, device=x.device)
    if n == 0:
        return y
    with torch.cuda.device(x.device):
        L.check(L.lib().stm_pool_fc_fwd(xn.data_ptr(), _dt(xn, "x"), n, h * w, c, xn.stride(0), xn.stride(3), wf.data_ptr(),
                                        bf.data_ptr() if bf is not None else None, wf.shape[0], y.data_ptr(), _stream(x)), "stm_pool_fc_fwd")
    return y


# --------------------------------------------------------------------------------------------
# torch.library registration (CUDA key only, fake impl for shape inference, no CPU key)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("stmask_b200::deform_conv2d", mutates_args=(), device_types="cuda")
def _deform_conv2d_op(x: torch.Tensor, offset: torch.Tensor, mask: Optional[torch.Tensor], weight: torch.Tensor,
                      bias: Optional[torch.Tensor], stride: List[int], padding: List[int], dilation: List[int],
                      groups: int, deform_groups: int, relu: bool, mask_sigmoid: bool) -> torch.Tensor:
    return deform_conv2d(x, offset, weight, bias, mask, stride, padding, dilation, groups, deform_groups, relu=relu,
                         mask_sigmoid=mask_sigmoid)


@_deform_conv2d_op.register_fake
def _(x, offset, mask, weight, bias, stride, padding, dilation, groups, deform_groups, relu, mask_sigmoid):
    spec = ConvSpec(x.shape[1], weight.shape[0], weight.shape[2:], stride, padding, dilation, groups, deform_groups)
    ho, wo = spec.out_hw(x.shape[2], x.shape[3])
    return torch.empty((x.shape[0], weight.shape[0], ho, wo), dtype=x.dtype, device=x.device,
                       memory_format=torch.channels_last)


@torch.library.custom_op("stmask_b200::correlation", mutates_args=(), device_types="cuda")
def _correlation_op(x1: torch.Tensor, x2: torch.Tensor, patch_size: int, dilation_patch: int, scale: float,
                    leaky_slope: float, relu: bool) -> torch.Tensor:
    return correlation(x1, x2, patch_size, dilation_patch, scale=scale,
                       leaky_slope=leaky_slope if leaky_slope != 0.0 else None, relu=relu)


@_correlation_op.register_fake
def _(x1, x2, patch_size, dilation_patch, scale, leaky_slope, relu):
    return x1.new_empty((x1.shape[0], patch_size * patch_size, x1.shape[2], x1.shape[3]))


def track_update(state: dict, dets: dict, mask_iou: torch.Tensor, is_first: Optional[torch.Tensor], *, match_coeff,
                 bbox_dummy_iou: float = 0.3, conf_thresh: float = 0.05, max_age: int = 10):
    """One frame of the tracker's matching state machine for a batch of clips, in place on the device state
    (Track_TF.track, track_TF.py:52-181; compute_comp_scores, TF_utils.py:98-123).

    state: dict of contiguous device tensors n_obj [C] int32, box [C, cap, 4], score [C, cap], cls [C, cap] int32,
    coeff [C, cap, k], track [C, cap, e], centerness [C, cap], tracked [C, cap] int32, mask_bits [C, cap, words] int32,
    mask [C, cap, h, w] (optional).  dets: count [C] int32 (optional), box / score / cls / coeff / track / centerness /
    mask_bits / mask with max_det rows per clip.  mask_iou [C, max_det, cap] (mask_iou_bits(det bits, state bits)).
    Returns (det_slot [C, max_det] int32, keep [C, cap] bool); nothing is synchronised."""
    box = state["box"]
    _require_cuda(box, "state box")
    if box.dim() != 3 or box.shape[2] != 4:
        raise ValueError("state box must be [clips, cap, 4]")
    clips, cap = box.shape[:2]
    dbox = dets["box"]
    if dbox.dim() != 3 or dbox.shape[0] != clips or dbox.shape[2] != 4:
        raise ValueError("detection box must be [clips, max_det, 4]")
    max_det = dbox.shape[1]
    k, e, words = state["coeff"].shape[2], state["track"].shape[2], state["mask_bits"].shape[2]
    if tuple(mask_iou.shape) != (clips, max_det, cap):
        raise ValueError(f"mask_iou must be {(clips, max_det, cap)}, got {tuple(mask_iou.shape)}")
    if len(match_coeff) != 4:
        raise ValueError("match_coeff needs 4 entries (score, mask IoU, box IoU, label)")
    want = {"n_obj": ((clips,), torch.int32), "box": ((clips, cap, 4), torch.float32), "score": ((clips, cap), torch.float32),
            "cls": ((clips, cap), torch.int32), "coeff": ((clips, cap, k), torch.float32), "track": ((clips, cap, e), torch.float32),
            "tracked": ((clips, cap), torch.int32), "mask_bits": ((clips, cap, words), torch.int32)}
    for name, (shape, dt) in want.items():
        t = state[name]
        if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.device != box.device:
            raise ValueError(f"state[{name!r}] must be a contiguous {dt} tensor of shape {shape} on {box.device}")
    dwant = {"box": ((clips, max_det, 4), torch.float32), "score": ((clips, max_det), torch.float32), "cls": ((clips, max_det), torch.int32),
             "coeff": ((clips, max_det, k), torch.float32), "track": ((clips, max_det, e), torch.float32),
             "mask_bits": ((clips, max_det, words), torch.int32)}
    for name, (shape, dt) in dwant.items():
        t = dets[name]
        if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.device != box.device:
            raise ValueError(f"dets[{name!r}] must be a contiguous {dt} tensor of shape {shape} on {box.device}")
    smask, dmask = state.get("mask"), dets.get("mask")
    hw = 0
    if smask is not None and dmask is not None:
        hw = smask[0, 0].numel()
        if smask.shape[:2] != (clips, cap) or dmask.shape[:2] != (clips, max_det) or dmask[0, 0].numel() != hw or \
                smask.dtype != torch.float32 or dmask.dtype != torch.float32 or not smask.is_contiguous() or not dmask.is_contiguous():
            raise ValueError("soft masks must be contiguous float32 [clips, cap, h, w] / [clips, max_det, h, w]")
    scent, dcent = state.get("centerness"), dets.get("centerness")
    for t, n, rows in ((scent, "state centerness", cap), (dcent, "dets centerness", max_det)):
        if t is not None and (tuple(t.shape) != (clips, rows) or t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError(f"{n} must be a contiguous float32 [clips, {rows}] tensor")
    count = dets.get("count")
    if count is not None and (count.dtype != torch.int32 or tuple(count.shape) != (clips,)):
        raise ValueError("dets count must be int32 [clips]")
    if is_first is not None:
        is_first = is_first.to(device=box.device, dtype=torch.uint8).contiguous()
        if tuple(is_first.shape) != (clips,):
            raise ValueError("is_first must be [clips]")
    miou = mask_iou.float().contiguous()
    det_slot = torch.full((clips, max_det), -1, dtype=torch.int32, device=box.device)
    keep = torch.zeros((clips, cap), dtype=torch.uint8, device=box.device)
    ptr = lambda t: t.data_ptr() if t is not None else None
    st = L.StmTrackState(ptr(state["n_obj"]), ptr(box), ptr(state["score"]), ptr(state["cls"]), ptr(state["coeff"]), ptr(state["track"]),
                         ptr(scent), ptr(state["tracked"]), ptr(state["mask_bits"]), ptr(smask) if hw else None)
    dt_ = L.StmTrackDets(ptr(count), ptr(dbox), ptr(dets["score"]), ptr(dets["cls"]), ptr(dets["coeff"]), ptr(dets["track"]),
                         ptr(dcent), ptr(dets["mask_bits"]), ptr(dmask) if hw else None)
    prm = L.StmTrackParams(clips, cap, max_det, k, e, words, hw, int(max_age), (C.c_float * 4)(*[float(v) for v in match_coeff]),
                           float(bbox_dummy_iou), float(conf_thresh))
    with torch.cuda.device(box.device):
        L.check(L.lib().stm_track_update_fwd(C.byref(prm), C.byref(st), C.byref(dt_), miou.data_ptr(), ptr(is_first), det_slot.data_ptr(),
                                             keep.data_ptr(), _stream(box)), "stm_track_update_fwd")
    return det_slot, keep.bool()
