"""ctypes loader for libedgeline_b200.so (the C ABI in include/edgeline_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libedgeline_b200.so")

EL_F32, EL_F16, EL_BF16 = 0, 1, 2
I64P = POINTER(c_int64)


class EdgelineError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "el_version": (c_char_p, []),
    "el_status_string": (c_char_p, [c_int]),
    "el_last_cuda_error": (c_int, []),
    "el_launch_count": (ctypes.c_uint64, []),
    "el_dwt_haar_fwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_dwt_haar_bwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_wave_merge_fwd": (c_int, [c_void_p, I64P, POINTER(c_void_p), I64P, c_void_p, c_void_p, I64P] + [c_int] * 7 + [c_void_p]),
    "el_wave_merge_bwd": (c_int, [c_void_p, I64P, POINTER(c_void_p), I64P, c_void_p, c_void_p, I64P, POINTER(c_void_p), I64P, c_void_p]
                          + [c_int] * 7 + [c_void_p]),
    "el_gated_residual_fwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_void_p, c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p]),
    "el_linattn_bwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_void_p]),
    "el_gated_residual_bwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_void_p, c_void_p, I64P, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p]),
    "el_linattn_fwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_void_p]),
    "el_gfl_decode_fwd": (c_int, [c_int, POINTER(c_void_p), I64P, POINTER(c_void_p), I64P, POINTER(c_int32), POINTER(c_float),
                                  POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_void_p), c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "el_gfl_detect_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "el_gfl_detect_fwd": (c_int, [c_int, POINTER(c_void_p), I64P, POINTER(c_void_p), I64P, POINTER(c_int32), POINTER(c_float),
                                  POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_void_p), c_int, c_int, c_int, c_float, c_double, c_int, c_int, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_size_t,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "el_nms_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "el_nms_batched": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_double, c_int, c_int, c_void_p, c_int, c_int, c_float,
                               c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "el_nms_boxes_workspace_bytes": (c_int, [c_int, POINTER(c_size_t)]),
    "el_nms_boxes": (c_int, [c_void_p, c_void_p, c_int, c_double, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "el_qfl_partials": (c_int, [c_int64]),
    "el_qfl_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "el_qfl_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "el_dfl_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "el_dfl_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "el_dfl_side_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "el_dfl_side_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "el_dsconv3_ok": (c_int, [c_int, c_int]),
    "el_dsconv3_fwd": (c_int, [c_void_p, I64P, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_conv3x3_mma_ok": (c_int, [c_int, c_int, c_int]),
    "el_conv3x3_mma_fwd": (c_int, [c_void_p, I64P, c_int, c_void_p, c_void_p, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_box_iou": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_float, c_void_p]),
    "el_ap_per_class_workspace_bytes": (c_int, [c_int64, c_int, c_int, POINTER(c_size_t)]),
    "el_ap_per_class": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_double, c_void_p, c_size_t, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "el_scale_boxes": (c_int, [c_void_p, c_int64, c_int64, c_float, c_float, c_float, c_int, c_int, c_float, c_float, c_void_p]),
    "el_match_predictions": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "el_tal_workspace_bytes": (c_int, [c_int, c_int, c_int, POINTER(c_size_t)]),
    "el_tal_assign": (c_int, [c_void_p] * 6 + [c_int] * 5 + [c_float] * 3 + [c_void_p, c_size_t] + [c_void_p] * 6),
    "el_ingest_u8": (c_int, [c_void_p, c_void_p, I64P, c_int, c_int, c_int, c_int, c_void_p]),
    "el_stem_conv_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_bias_act_fwd": (c_int, [c_void_p, I64P, c_void_p, c_void_p, I64P, c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_void_p]),
    "el_dwconv_fwd": (c_int, [c_void_p, I64P, c_void_p, c_void_p, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_pwconv_tile": (c_int, [c_int, c_int, c_int64]),
    "el_pwconv_fwd": (c_int, [c_int, POINTER(c_void_p), I64P, POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_int64, c_float, c_int, c_int, c_void_p, c_int64, c_void_p,
                              c_int64, c_int, c_int64, c_int, c_int, c_int, c_void_p]),
    "el_conv3x3_tile": (c_int, [c_int, c_int, c_int64]),
    "el_conv3x3_fwd": (c_int, [c_void_p, I64P, c_int, c_void_p, c_void_p, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_conv3x3_halo_ok": (c_int, [c_int, c_int]),
    "el_conv3x3_halo_fwd": (c_int, [c_void_p, I64P, c_int, c_void_p, c_void_p, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_sppf_pool_fwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "el_upsample2x_cat_fwd": (c_int, [c_void_p, I64P, c_void_p, I64P, c_void_p, I64P, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}

EXPORTED = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises EdgelineError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EdgelineError(
                f"{LIB_PATH} is missing: build it with `python -m edge_yolo_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here means header and library disagree
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        L = lib()
        msg = L.el_status_string(status).decode()
        if status == 4:
            msg += f" (cudaError {L.el_last_cuda_error()})"
        raise EdgelineError(f"{what}: {msg}")
