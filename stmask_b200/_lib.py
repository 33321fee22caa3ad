"""ctypes binding of libstmask_b200.so (include/stmask_b200.h).

The product path has NO CPU implementation: if the shared library is missing or fails to load
this module raises — it never falls back to eager PyTorch or to the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstmask_b200.so")

STM_OK = 0
STM_F32, STM_BF16 = 0, 1
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05 = 0, 1, 2
BACKEND_NAMES = {BACKEND_AUTO: "auto", BACKEND_SIMT: "simt", BACKEND_TCGEN05: "tcgen05"}
DCN_RELU, DCN_MASK_SIGMOID, DCN_ZERO_OFFSET = 1, 2, 4
DCN_HINT_RASTER = 8
DCN_HINT_ROWS128, DCN_HINT_ROWS256, DCN_HINT_NO_PAIR = 16, 32, 64
DCN_OUT_F32, DCN_HINT_DEEP_PIPE, DCN_HINT_TWO_CTAS = 128, 256, 512
DCN_OUT_PLANAR = 131072
DCN_FCB_ADA, DCN_FCB_ALI = 1024, 2048
DCN_HINT_GATHER = 4096
DCN_HINT_TAP_MAJOR, DCN_HINT_CHUNK_MAJOR, DCN_HINT_NO_FUSE = 8192, 16384, 32768
CORR_LEAKY_RELU, CORR_RELU, CORR_COPY_FEATS = 1, 2, 4
DCN_MAX_PROBLEMS = 8
ABI_VERSION = 6


class StmError(RuntimeError):
    """A call into libstmask_b200.so returned a non-zero status."""


class StmDcnConv(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "in_c", "out_c", "kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w",
        "groups", "deform_groups", "dtype", "offset_dtype", "flags", "backend")]


class StmDcnProblem(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("x", C.c_void_p), ("x_stride_n", C.c_int64), ("x_stride_h", C.c_int64), ("x_stride_w", C.c_int64),
        ("offset", C.c_void_p), ("off_stride_n", C.c_int64), ("off_stride_c", C.c_int64),
        ("off_stride_h", C.c_int64), ("off_stride_w", C.c_int64),
        ("mask", C.c_void_p), ("mask_stride_n", C.c_int64), ("mask_stride_c", C.c_int64),
        ("mask_stride_h", C.c_int64), ("mask_stride_w", C.c_int64),
        ("y", C.c_void_p), ("y_stride_n", C.c_int64), ("y_stride_h", C.c_int64), ("y_stride_w", C.c_int64),
    ]


class StmCorrDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("patch", C.c_int32), ("dilation_patch", C.c_int32), ("dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("flags", C.c_int32), ("backend", C.c_int32), ("scale", C.c_float), ("leaky_slope", C.c_float),
        ("x1_stride_n", C.c_int64), ("x1_stride_h", C.c_int64), ("x1_stride_w", C.c_int64),
        ("x2_stride_n", C.c_int64), ("x2_stride_h", C.c_int64), ("x2_stride_w", C.c_int64),
        ("out_stride_n", C.c_int64), ("out_stride_c", C.c_int64), ("out_stride_h", C.c_int64), ("out_stride_w", C.c_int64),
        ("feat_c", C.c_int32), ("feat_dtype", C.c_int32),
        ("feat_a_stride_n", C.c_int64), ("feat_a_stride_h", C.c_int64), ("feat_a_stride_w", C.c_int64),
        ("feat_b_stride_n", C.c_int64), ("feat_b_stride_h", C.c_int64), ("feat_b_stride_w", C.c_int64),
        ("feat_c_offset", C.c_int32), ("reserved_", C.c_int32),
        ("x1_index", C.c_void_p), ("x2_index", C.c_void_p),
        ("x1_frames", C.c_int32), ("x2_frames", C.c_int32), ("alt_frames", C.c_int32), ("reserved2_", C.c_int32),
        ("x1_alt", C.c_void_p), ("feat_a_alt", C.c_void_p),
        ("x1_alt_stride_n", C.c_int64), ("x1_alt_stride_h", C.c_int64), ("x1_alt_stride_w", C.c_int64),
        ("feat_a_alt_stride_n", C.c_int64), ("feat_a_alt_stride_h", C.c_int64), ("feat_a_alt_stride_w", C.c_int64),
    ]


class StmRoiAlignDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("n_rois", C.c_int32),
        ("pooled_h", C.c_int32), ("pooled_w", C.c_int32), ("sampling_ratio", C.c_int32), ("aligned", C.c_int32),
        ("dtype", C.c_int32), ("out_dtype", C.c_int32), ("spatial_scale", C.c_float),
        ("feat_stride_n", C.c_int64), ("feat_stride_h", C.c_int64), ("feat_stride_w", C.c_int64),
        ("out_stride_n", C.c_int64), ("out_stride_c", C.c_int64), ("out_stride_h", C.c_int64), ("out_stride_w", C.c_int64),
    ]


class StmTrackState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("n_obj", "box", "score", "cls", "coeff", "track", "centerness", "tracked", "mask_bits", "mask")]


class StmTrackDets(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("count", "box", "score", "cls", "coeff", "track", "centerness", "mask_bits", "mask")]


class StmTrackParams(C.Structure):
    _fields_ = [("clips", C.c_int32), ("cap", C.c_int32), ("max_det", C.c_int32), ("k", C.c_int32), ("e", C.c_int32),
                ("words", C.c_int32), ("hw", C.c_int32), ("max_age", C.c_int32), ("match_coeff", C.c_float * 4),
                ("bbox_dummy_iou", C.c_float), ("conf_thresh", C.c_float)]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "stm_version": (C.c_int, []),
    "stm_last_error": (C.c_char_p, []),
    "stm_device_supported": (C.c_int, [C.c_int32]),
    "stm_kernel_launch_count": (C.c_uint64, []),
    "stm_dcn_packed_weight_bytes": (C.c_size_t, [C.POINTER(StmDcnConv)]),
    "stm_dcn_pack_weight": (C.c_int, [C.POINTER(StmDcnConv), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stm_deform_conv2d_workspace": (C.c_size_t, [C.POINTER(StmDcnConv), C.POINTER(StmDcnProblem), C.c_int32]),
    "stm_deform_conv2d_fwd": (C.c_int, [C.POINTER(StmDcnConv), C.POINTER(StmDcnProblem), C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stm_deform_conv2d_fcb_fwd": (C.c_int, [C.POINTER(StmDcnConv), C.POINTER(StmDcnProblem), C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "stm_deform_conv2d_backend": (C.c_int, [C.POINTER(StmDcnConv), C.POINTER(StmDcnProblem), C.c_int32]),
    "stm_deform_conv2d_variant": (C.c_int, [C.POINTER(StmDcnConv), C.POINTER(StmDcnProblem), C.c_int32, C.c_char_p, C.c_size_t]),
    "stm_fcb_ali_offsets": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.POINTER(C.c_int64),
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "stm_fcb_ada_offsets": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p]),
    "stm_correlation_fwd": (C.c_int, [C.POINTER(StmCorrDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "stm_correlation_backend": (C.c_int, [C.POINTER(StmCorrDesc)]),
    "stm_correlation_multi_fwd": (C.c_int, [C.POINTER(StmCorrDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "stm_roi_align_fwd": (C.c_int, [C.POINTER(StmRoiAlignDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stm_detect_fast_nms_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stm_mask_assembly_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "stm_mask_iou_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]),
    "stm_track_update_fwd": (C.c_int, [C.POINTER(StmTrackParams), C.POINTER(StmTrackState), C.POINTER(StmTrackDets), C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stm_pool_fc_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stm_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]),
    "stm_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p]),
}

_LIB: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise StmError(
                f"{LIB_PATH} is missing. stmask_b200 has no CPU or eager fallback: build the CUDA library first "
                f"(`python -m stmask_b200.build`, or `__graft_entry__.build()`).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the export is missing
            fn.restype, fn.argtypes = res, args
        v = handle.stm_version()
        if v != ABI_VERSION:
            raise StmError(f"libstmask_b200.so ABI version {v}, Python binding expects {ABI_VERSION}")
        _LIB = handle
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != STM_OK:
        msg = lib().stm_last_error().decode(errors="replace")
        raise StmError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(lib().stm_kernel_launch_count())
