"""ctypes front-end of ``stm_oracle.c`` plus a tiny pure-numpy restatement.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  All functions take and return
contiguous float32 numpy arrays in the reference's NCHW layout.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Sequence, Tuple, Union

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[ctypes.CDLL] = None

IntPair = Union[int, Sequence[int]]


def lib_path() -> str:
    return os.path.join(_HERE, "_build", "libstm_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc, OpenMP)."""
    path = lib_path()
    src = os.path.join(_HERE, "stm_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return path


def _lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(lib_path()):
            build()
        lib = ctypes.CDLL(lib_path())
        f32p = ctypes.POINTER(ctypes.c_float)
        lib.stm_oracle_deform_conv2d.restype = ctypes.c_int
        lib.stm_oracle_deform_conv2d.argtypes = [f32p] * 6 + [ctypes.c_int] * 16
        lib.stm_oracle_correlation.restype = ctypes.c_int
        lib.stm_oracle_correlation.argtypes = [f32p] * 3 + [ctypes.c_int] * 7
        lib.stm_oracle_correlate_post.restype = None
        lib.stm_oracle_correlate_post.argtypes = [f32p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float]
        lib.stm_oracle_fcb_ali_offsets.restype = None
        lib.stm_oracle_fcb_ali_offsets.argtypes = [f32p, f32p] + [ctypes.c_int] * 5
        lib.stm_oracle_roi_align.restype = ctypes.c_int
        lib.stm_oracle_roi_align.argtypes = [f32p] * 3 + [ctypes.c_int] * 7 + [ctypes.c_float, ctypes.c_int, ctypes.c_int]
        lib.stm_oracle_num_threads.restype = ctypes.c_int
        lib.stm_oracle_out_size.restype = ctypes.c_int
        lib.stm_oracle_out_size.argtypes = [ctypes.c_int] * 5
        _LIB = lib
    return _LIB


def _pair(v: IntPair) -> Tuple[int, int]:
    if isinstance(v, int):
        return v, v
    a, b = v
    return int(a), int(b)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return ctypes.POINTER(ctypes.c_float)()
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def num_threads() -> int:
    return int(_lib().stm_oracle_num_threads())


def out_size(n: int, k: int, s: int, p: int, d: int) -> int:
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


def deform_conv2d(x, offset, weight, bias=None, mask=None, stride: IntPair = 1, padding: IntPair = 0,
                  dilation: IntPair = 1, groups: int = 1, deform_groups: int = 1,
                  accum64: bool = True) -> np.ndarray:
    """DCNv1 (mask None) / DCNv2 forward.  Restates dcn_v2_conv / mmcv deform_conv2d as
    called from reference backbone.py:45 and Featurealign.py:72."""
    x, weight = _f32(x), _f32(weight)
    offset = None if offset is None else _f32(offset)
    mask = None if mask is None else _f32(mask)
    bias = None if bias is None else _f32(bias)
    B, Cin, H, W = x.shape
    Cout, cpg, kh, kw = weight.shape
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    if cpg * groups != Cin:
        raise ValueError("weight/in_channels/groups mismatch")
    Ho, Wo = out_size(H, kh, sh, ph, dh), out_size(W, kw, sw, pw, dw)
    if offset is not None and offset.shape != (B, deform_groups * 2 * kh * kw, Ho, Wo):
        raise ValueError(f"offset shape {offset.shape} != {(B, deform_groups * 2 * kh * kw, Ho, Wo)}")
    if mask is not None and mask.shape != (B, deform_groups * kh * kw, Ho, Wo):
        raise ValueError(f"mask shape {mask.shape}")
    y = np.empty((B, Cout, Ho, Wo), np.float32)
    rc = _lib().stm_oracle_deform_conv2d(_ptr(x), _ptr(offset), _ptr(mask), _ptr(weight), _ptr(bias), _ptr(y),
                                         B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw,
                                         groups, deform_groups, int(accum64))
    if rc != 0:
        raise ValueError(f"stm_oracle_deform_conv2d failed: {rc}")
    return y


def correlation(x1, x2, patch_size: int = 11, dilation_patch: int = 1, accum64: bool = True) -> np.ndarray:
    """spatial_correlation_sample(kernel_size=1, stride=1, padding=0) -> [B, P, P, H, W]
    (reference track_to_segment_head.py:53-59)."""
    x1, x2 = _f32(x1), _f32(x2)
    if x1.shape != x2.shape:
        raise ValueError("x1/x2 shape mismatch")
    B, C, H, W = x1.shape
    out = np.empty((B, patch_size, patch_size, H, W), np.float32)
    rc = _lib().stm_oracle_correlation(_ptr(x1), _ptr(x2), _ptr(out), B, C, H, W,
                                       patch_size, dilation_patch, int(accum64))
    if rc != 0:
        raise ValueError(f"stm_oracle_correlation failed: {rc}")
    return out


def correlate(x1, x2, patch_size: int = 11, dilation_patch: int = 1, accum64: bool = True) -> np.ndarray:
    """The reference's ``correlate`` (track_to_segment_head.py:40-62): cost volume viewed as
    [B, P*P, H, W], divided by C, leaky-ReLU(0.1)."""
    out = correlation(x1, x2, patch_size, dilation_patch, accum64)
    B, P, _, H, W = out.shape
    out = out.reshape(B, P * P, H, W)
    _lib().stm_oracle_correlate_post(_ptr(out), out.size, int(np.asarray(x1).shape[1]), 0.1)
    return out


def roi_align(feat, rois, output_size: IntPair = 7, spatial_scale: float = 1.0, sampling_ratio: int = 0,
              aligned: bool = True) -> np.ndarray:
    """mmcv.ops.roi_align (avg) as the reference calls it (track_to_segment_head.py:85-86): feat [B, C, H, W],
    rois [n, 5] = (batch index, x1, y1, x2, y2) -> [n, C, ph, pw]."""
    feat, rois = _f32(feat), _f32(rois)
    ph, pw = _pair(output_size)
    B, C, H, W = feat.shape
    out = np.empty((rois.shape[0], C, ph, pw), np.float32)
    rc = _lib().stm_oracle_roi_align(_ptr(feat), _ptr(rois), _ptr(out), B, C, H, W, rois.shape[0], ph, pw,
                                     float(spatial_scale), int(sampling_ratio), int(bool(aligned)))
    if rc != 0:
        raise ValueError("stm_oracle_roi_align: bad arguments")
    return out


def fcb_ali_offsets(shape, kernel_size: IntPair) -> np.ndarray:
    """Closed-form FCB(ali) offsets (Featurealign.py:46-69)."""
    shape = _f32(shape)
    kh, kw = _pair(kernel_size)
    B, four, H, W = shape.shape
    assert four == 4
    off = np.empty((B, 2 * kh * kw, H, W), np.float32)
    _lib().stm_oracle_fcb_ali_offsets(_ptr(shape), _ptr(off), B, H, W, kh, kw)
    return off


def feature_align(x, shape, w_adaption, kernel_size: IntPair, w_offset=None, deform_groups: int = 1,
                  accum64: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """FeatureAlign.forward up to and including the ReLU (Featurealign.py:42-72).

    ``w_offset`` (the 1x1 ``conv_offset`` weight [dg*2*kh*kw, 4, 1, 1]) selects FCB(ada);
    ``None`` selects FCB(ali).  Returns (relu(deform_conv(x, offset)), offset)."""
    kh, kw = _pair(kernel_size)
    shape = _f32(shape)
    if w_offset is not None:
        wo = _f32(w_offset).reshape(-1, 4).astype(np.float64)
        offset = np.einsum("oc,bchw->bohw", wo, shape.astype(np.float64)).astype(np.float32)
    else:
        offset = fcb_ali_offsets(shape, (kh, kw))
    y = deform_conv2d(x, offset, w_adaption, padding=((kh - 1) // 2, (kw - 1) // 2),
                      deform_groups=deform_groups, accum64=accum64)
    return np.maximum(y, 0.0), offset


# --------------------------------------------------------------------------------------
# pure-numpy restatement (slow; small cases only) — an independent check of the C code
# --------------------------------------------------------------------------------------
def _np_bilinear(plane: np.ndarray, h: float, w: float) -> float:
    H, W = plane.shape
    if h <= -1 or h >= H or w <= -1 or w >= W:
        return 0.0
    h0, w0 = math.floor(h), math.floor(w)
    lh, lw = h - h0, w - w0
    val = 0.0
    for (yy, xx, wt) in ((h0, w0, (1 - lh) * (1 - lw)), (h0, w0 + 1, (1 - lh) * lw),
                         (h0 + 1, w0, lh * (1 - lw)), (h0 + 1, w0 + 1, lh * lw)):
        if 0 <= yy < H and 0 <= xx < W:
            val += wt * float(plane[yy, xx])
    return val


def np_deform_conv2d(x, offset, weight, bias=None, mask=None, stride=1, padding=0, dilation=1,
                     groups=1, deform_groups=1) -> np.ndarray:
    x = np.asarray(x, np.float64)
    weight = np.asarray(weight, np.float64)
    B, Cin, H, W = x.shape
    Cout, cpg, kh, kw = weight.shape
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    Ho, Wo = out_size(H, kh, sh, ph, dh), out_size(W, kw, sw, pw, dw)
    K = kh * kw
    cpd = Cin // deform_groups
    opg = Cout // groups
    y = np.zeros((B, Cout, Ho, Wo), np.float64)
    for b in range(B):
        col = np.zeros((Cin, K, Ho, Wo), np.float64)
        for c in range(Cin):
            g = c // cpd
            for i in range(kh):
                for j in range(kw):
                    k = i * kw + j
                    for ho in range(Ho):
                        for wo in range(Wo):
                            oy = ox = 0.0
                            if offset is not None:
                                oy = float(offset[b, g * 2 * K + 2 * k, ho, wo])
                                ox = float(offset[b, g * 2 * K + 2 * k + 1, ho, wo])
                            m = 1.0 if mask is None else float(mask[b, g * K + k, ho, wo])
                            col[c, k, ho, wo] = m * _np_bilinear(
                                x[b, c], ho * sh - ph + i * dh + oy, wo * sw - pw + j * dw + ox)
        for co in range(Cout):
            g = co // opg
            y[b, co] = np.tensordot(weight[co].reshape(cpg, K), col[g * cpg:(g + 1) * cpg], axes=([0, 1], [0, 1]))
            if bias is not None:
                y[b, co] += float(bias[co])
    return y.astype(np.float32)


def np_correlation(x1, x2, patch_size=11, dilation_patch=1) -> np.ndarray:
    """Shifted-product form of the definition."""
    x1 = np.asarray(x1, np.float64)
    x2 = np.asarray(x2, np.float64)
    B, C, H, W = x1.shape
    P, d = patch_size, dilation_patch
    r = P // 2
    pad = r * d
    x2p = np.pad(x2, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    out = np.zeros((B, P, P, H, W), np.float64)
    for ph in range(P):
        for pw in range(P):
            dy, dx = (ph - r) * d + pad, (pw - r) * d + pad
            out[:, ph, pw] = (x1 * x2p[:, :, dy:dy + H, dx:dx + W]).sum(1)
    return out.astype(np.float32)
