"""ctypes binding of libebos.so (the C-ABI declared in include/ebos.h).

This is the only place the package touches native code.  There is no fallback: if the library is
missing it is built with nvcc (event_based_bos_b200/_build.py); if that fails, or no CUDA device is
present when a kernel is needed, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

from . import _build

# error codes / enums (mirror include/ebos.h)
EBOS_OK = 0
EBOS_F32, EBOS_F64 = 0, 1
DIR_FIRST, DIR_LAST, DIR_FRAC = 0, 1, 2
COST_NONE, COST_VARIANCE, COST_GRADMAG = 0, 1, 2
ACC_DOUBLES = 40  # EBOS_ACC_DOUBLES (include/ebos.h)
STATUS_PIXEL_OOB = 1
STATUS_PACKED = 2
WIN_HAS_WEIGHT, WIN_PACKED = 1, 2
EKLT_POISSON, EKLT_WARP, EKLT_NO_POLARITY = 1, 2, 4

# name -> (restype, argtypes); every symbol include/ebos.h declares.
SIGNATURES = {
    "ebos_version": (c_int, []),
    "ebos_last_error": (c_char_p, []),
    "ebos_device_available": (c_int, []),
    "ebos_time_stats": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "ebos_warp_dense_flow": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_void_p, c_int,
                                     c_double, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ebos_warp_dense_flow_bwd": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_double, c_int,
                                         c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "ebos_warp_2dof": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_double, c_int, c_int,
                               c_void_p, c_void_p]),
    "ebos_iwe_splat": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_double, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ebos_splat_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int, c_int]),
    "ebos_iwe_splat_bwd": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_double, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "ebos_window_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "ebos_window_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "ebos_window_prepare": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p, c_int,
                                    c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "ebos_window_info": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ebos_window_splat": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p]),
    "ebos_iwe_cost": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "ebos_flow_tv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "ebos_window_backward": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                     c_int, c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "ebos_loss_finalize": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_int,
                                   c_void_p, c_void_p]),
    "ebos_cmax_value_and_grad": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_double, c_double, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "ebos_cmax_adam_iteration": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_double, c_double, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_double, c_void_p,
                                         c_double, c_void_p, c_void_p]),
    "ebos_cmax_adam_iteration_fused_tv": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                  c_int, c_int, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double,
                                                  c_double, c_void_p, c_double, c_void_p, c_void_p]),
    "ebos_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double,
                               c_int, c_int, c_void_p]),
    "ebos_iwe_cost_peers": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "ebos_sum_peers": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "ebos_multimem_allreduce_slice": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "ebos_reduce_peers_slice": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "ebos_gather_peers_slices": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "ebos_ingest_workspace_bytes": (c_size_t, [c_int64]),
    "ebos_ingest_raw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int64,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ebos_time_to_index": (c_int, [c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "ebos_adam_step_graph": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double,
                                     c_double, c_void_p, c_int, c_void_p]),
    "ebos_eklt_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "ebos_eklt_value_and_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_int,
                                         c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ebos_eklt_adam_iteration": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_int,
                                         c_void_p, c_size_t,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_double,
                                         c_void_p, c_void_p]),
    "ebos_eklt_upsample": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ebos_eklt_patch_flow": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ebos_sepconv2d": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p]),
    "ebos_flow_error_workspace_doubles": (c_size_t, [c_int]),
    "ebos_flow_error": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                c_void_p, c_void_p]),
    "ebos_blur3": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_void_p, c_void_p]),
    "ebos_capture_begin": (c_int, [c_void_p]),
    "ebos_capture_end_count": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ebos_capture_end_exec": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ebos_exec_launch": (c_int, [c_void_p, c_void_p]),
    "ebos_exec_destroy": (c_int, [c_void_p]),
}

_lib = None
_lock = threading.Lock()


def library_path() -> str:
    return os.environ.get("EBOS_LIBRARY", _build.LIB_PATH)


def load() -> ctypes.CDLL:
    """Load (building first if needed) libebos.so and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if path == _build.LIB_PATH and not os.path.exists(path):
            _build.build_library()
        if not os.path.exists(path):
            raise RuntimeError(f"libebos.so not found at {path}; build it with `python -m event_based_bos_b200._build`")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().ebos_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != EBOS_OK:
        raise RuntimeError(f"libebos {what} failed (code {rc}): {last_error()}")


def require_device() -> None:
    """Fail loudly when the CUDA path cannot run (no silent CPU fallback exists)."""
    import torch

    if not torch.cuda.is_available() or not load().ebos_device_available():
        raise RuntimeError(
            "event_based_bos_b200 needs a CUDA device (B200, sm_100a): its operators run only as CUDA kernels "
            "and there is no CPU implementation to fall back to.")


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()
