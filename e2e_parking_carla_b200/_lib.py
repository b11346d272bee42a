"""ctypes binding of the C ABI declared in ``include/ls_b200.h``.

The library is the product: there is no CPU or eager-PyTorch fallback.  Loading
fails loudly when ``libls_b200.so`` is missing (build it with
``python -m e2e_parking_carla_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

LS_OK = 0
LS_F32, LS_BF16 = 0, 1
LS_FEAT_NCHW, LS_FEAT_NHWC = 0, 1
LS_GEOM_TORCH_CPU, LS_GEOM_TORCH_CUDA = 0, 1


class LsShape(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("D", C.c_int32), ("fh", C.c_int32),
                ("fw", C.c_int32), ("C", C.c_int32), ("X", C.c_int32), ("Y", C.c_int32),
                ("Z", C.c_int32), ("start", C.c_float * 3), ("res", C.c_float * 3),
                ("geom_policy", C.c_int32), ("tile_x", C.c_int32),
                ("bev_dtype", C.c_int32)]


class LsBevStrides(C.Structure):
    _fields_ = [("b", C.c_int64), ("c", C.c_int64), ("x", C.c_int64), ("y", C.c_int64)]


_P = C.c_void_p
_SH = C.POINTER(LsShape)
_ST = C.POINTER(LsBevStrides)

# name -> (restype, argtypes); must list every symbol of include/ls_b200.h
PROTOTYPES = {
    "ls_version": (C.c_char_p, []),
    "ls_strerror": (C.c_char_p, [C.c_int]),
    "ls_last_cuda_error": (C.c_char_p, []),
    "ls_launch_count": (C.c_int64, []),
    "ls_debug_phase_cycles": (C.c_int, [C.POINTER(C.c_uint64)]),
    "ls_grid_cells": (C.c_int, [_SH, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ls_padded_channels": (C.c_int32, [C.c_int32]),
    "ls_sorted_records": (C.c_size_t, [_SH]),
    "ls_camera_transform": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P]),
    "ls_geometry": (C.c_int, [_P, _P, _P, _SH, _P, _P]),
    "ls_index": (C.c_int, [_P, _P, _P, _SH, _P, _P, _P, _P, _P]),
    "ls_index_geom": (C.c_int, [_P, _SH, _P, _P, _P, _P, _P]),
    "ls_export_indices": (C.c_int, [_P, _P, _P, _SH, _P, _P, _P, _P]),
    "ls_sort": (C.c_int, [_P, _P, _P, _P, C.c_int, _SH, _P, _P, _P, _P, _P, _P]),
    "ls_export_cell_counts": (C.c_int, [_P, _SH, C.c_int32, _P, _P, _P]),
    "ls_softmax": (C.c_int, [_P, C.c_int, _SH, _P, _P]),
    "ls_nchw_to_nhwc": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ls_nhwc_to_nchw": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ls_splat_fwd": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _SH, _P, _ST, _P]),
    "ls_splat_bwd": (C.c_int, [_P, _ST, _P, C.c_int, _P, _P, _SH, _P, _P, _P, _P]),
    "ls_softmax_bwd": (C.c_int, [_P, _P, _P, C.c_int, _SH, _P, _P]),
    "ls_target_bev": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, C.c_int64, C.c_int64, _P]),
    "ls_depth_loss_ws_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "ls_depth_loss_fwd": (C.c_int, [_P, C.c_int, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_float, _P, _P, C.c_size_t, _P, _P]),
    "ls_depth_loss_bwd": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ls_scratch_bytes": (C.c_size_t, [_SH, C.c_int, C.c_int]),
    "ls_saved_bytes": (C.c_size_t, [_SH, C.c_int, C.c_int]),
    "ls_cache_bytes": (C.c_size_t, [_SH]),
    "ls_forward_cached": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _P, _SH, _P, C.c_size_t, _P, C.c_size_t, _P,
                                    C.c_size_t, C.c_int, _P, _ST, _P, _P]),
    "ls_backward_cached": (C.c_int, [_P, _ST, _P, _P, _P, C.c_int, C.c_int, _SH, _P, C.c_size_t, _P, C.c_size_t, _P,
                                     C.c_size_t, _P, _P, _P]),
    "ls_forward": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _P, _SH, _P, C.c_size_t, _P, C.c_size_t, _P, _ST, _P,
                             _P]),
    "ls_backward": (C.c_int, [_P, _ST, _P, _P, _P, C.c_int, C.c_int, _SH, _P, C.c_size_t, _P, C.c_size_t, _P, _P,
                              _P]),
}


class LiftSplatLibraryError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and attach prototypes.  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LiftSplatLibraryError(
            "CUDA library %s is missing - run `python -m e2e_parking_carla_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    """Turn an LsStatus into a Python exception (error behaviour of the boundary)."""
    if status == LS_OK:
        return
    lib = load()
    msg = lib.ls_strerror(status).decode()
    if status == -4:
        msg += ": " + lib.ls_last_cuda_error().decode()
    if status in (-1, -2):
        raise ValueError("%s: %s" % (what, msg))
    raise LiftSplatLibraryError("%s: %s" % (what, msg))
