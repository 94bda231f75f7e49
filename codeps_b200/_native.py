"""ctypes binding of libcodeps_photo.so (include/codeps_photo.h).

This is the only place the package touches native code.  There is no fallback: if the shared
library cannot be loaded (and cannot be built with nvcc), importing an operator raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_float, c_int32, c_size_t, c_uint8, c_uint64, c_void_p

MAX_LEVELS = 6
MAX_BATCH_PER_LAUNCH = 32
MAX_FLOW_MAPS = 4    # CDP_MAX_FLOW_MAPS
ABI_VERSION = 6

_fp = POINTER(c_float)


class PhotoHeads(ctypes.Structure):
    """struct cdp_photo_heads"""
    _fields_ = [
        ("disp", c_void_p), ("min_depth", c_float), ("max_depth", c_float),
        ("axisangle", c_void_p * 2), ("translation", c_void_p * 2), ("invert", c_int32 * 2),
    ]


class PhotoArgs(ctypes.Structure):
    """struct cdp_photo_args"""
    _fields_ = [
        ("batch", c_int32), ("height", c_int32), ("width", c_int32), ("num_levels", c_int32),
        ("alpha", c_float), ("with_grad", c_int32),
        ("intrinsics_host", c_void_p),
        ("target", c_void_p), ("source0", c_void_p), ("source1", c_void_p), ("depth", c_void_p),
        ("pose0", c_void_p), ("pose1", c_void_p),
        ("noise", c_void_p * MAX_LEVELS),
        ("noise_seed", c_uint64),
        ("resize_tables", c_void_p),
        ("loss", c_void_p),
        ("argmin", c_void_p * MAX_LEVELS),
        ("scratch", c_void_p), ("scratch_bytes", c_size_t),
        ("saved", c_void_p), ("saved_bytes", c_size_t),
        ("motion0", c_void_p), ("motion1", c_void_p),
        ("intrinsics_dev", c_void_p),
        ("noise_ready", c_void_p),
        ("heads", POINTER(PhotoHeads)),
        ("noise_seed_dev", c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/codeps_photo.h declares
SIGNATURES = {
    "cdp_version": (c_int32, []),
    "cdp_last_error": (c_char_p, []),
    "cdp_device_check": (c_int32, []),
    "cdp_profile_enable": (c_int32, [c_int32]),
    "cdp_profile_read": (c_int32, [c_int32, POINTER(ctypes.c_double), POINTER(c_int32)]),
    "cdp_resize_tables_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "cdp_resize_tables_build": (c_int32, [c_int32, c_int32, c_int32, c_void_p, c_size_t]),
    "cdp_photo_scratch_bytes": (c_size_t, [c_int32] * 5),
    "cdp_photo_saved_bytes": (c_size_t, [c_int32] * 5),
    "cdp_photo_fwd": (c_int32, [POINTER(PhotoArgs), c_void_p]),
    "cdp_photo_bwd": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "cdp_photo_bwd_heads": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p, c_void_p,
                                      POINTER(PhotoHeads), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int32, c_void_p, c_void_p, c_void_p]),
    "cdp_tiebreak_noise": (c_int32, [c_int32, c_int32, c_int32, c_int32, ctypes.c_uint64, c_void_p, c_void_p]),
    "cdp_photo_fwd_launches": (c_int32, [c_int32, c_int32]),
    "cdp_photo_bwd_launches": (c_int32, [c_int32, c_int32, c_int32]),
    "cdp_smooth_saved_bytes": (c_size_t, [c_int32] * 3),
    "cdp_smooth_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "cdp_smooth_bwd": (c_int32, [c_void_p, c_size_t, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                 c_void_p]),
    "cdp_warp_grid_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                    c_int32, c_void_p, c_void_p]),
    "cdp_warp_image_fwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "cdp_warp_bwd_scratch_bytes": (c_size_t, [c_int32] * 3),
    "cdp_warp_image_bwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_ssim_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "cdp_ssim_bwd_scratch_bytes": (c_size_t, [c_int32] * 3),
    "cdp_ssim_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_pose_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "cdp_pose_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "cdp_disp_to_depth_fwd": (c_int32, [c_void_p, c_size_t, c_float, c_float, c_void_p, c_void_p]),
    "cdp_disp_to_depth_bwd": (c_int32, [c_void_p, c_void_p, c_size_t, c_float, c_float, c_void_p, c_void_p]),
    "cdp_flow_scratch_bytes": (c_size_t, [c_int32] * 5),
    "cdp_flow_smooth_fwd": (c_int32, [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_flow_sparsity_fwd": (c_int32, [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int32, c_void_p,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_scale_fwd": (c_int32, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "cdp_depth_metrics_scratch_bytes": (c_size_t, [c_int32, c_int32]),
    "cdp_depth_metrics_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_int32, c_int32, c_int32, c_int32,
                                        c_int32, c_float, c_float, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdp_warp_c2c_fwd": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                   c_void_p, ctypes.c_double, c_int32, c_int32, c_void_p, c_void_p]),
}

_lock = threading.Lock()
_lib = None


class NativeError(RuntimeError):
    """A libcodeps_photo.so entry point returned a negative status."""


def library_path() -> str:
    # CODEPS_B200_LIB selects an alternative build of the same library (kernel tuning experiments)
    override = os.environ.get("CODEPS_B200_LIB")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcodeps_photo.so")


def load() -> ctypes.CDLL:
    """Load (building first if necessary) the CUDA library.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.environ.get("CODEPS_B200_LIB"):
            # (re)build when the library is missing or older than its sources, so that a kernel fix
            # that does not bump CDP_ABI_VERSION is never shadowed by a stale git-ignored .so
            from . import build
            missing = not os.path.exists(path)
            try:
                if missing or build.is_stale():
                    build.build_native()
            except Exception as exc:  # no nvcc, compile error ...
                if missing:
                    raise RuntimeError(
                        f"codeps_b200: {path} is missing and could not be built ({exc}). The package "
                        "has no CPU or PyTorch fallback; run `python -m codeps_b200.build`.") from exc
                import warnings
                warnings.warn(f"codeps_b200: {path} is older than its sources and could not be rebuilt ({exc}); "
                              "loading the stale library")
        elif not os.path.exists(path):
            raise RuntimeError(f"codeps_b200: CODEPS_B200_LIB={path} does not exist")
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise RuntimeError(f"codeps_b200: {path} does not export {name}; rebuild it") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.cdp_version() != ABI_VERSION:
            raise RuntimeError(f"codeps_b200: ABI version mismatch ({lib.cdp_version()} != "
                               f"{ABI_VERSION}); rebuild with `python -m codeps_b200.build --force`")
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().cdp_last_error()
        raise NativeError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


KERNEL_IDS = {"pyramid": 0, "photo": 1, "finalize": 2, "depth_grad": 3, "smooth_sum": 4,
              "smooth_main": 5, "smooth_finalize": 6, "smooth_bwd": 7}


def profile_enable(enable: bool) -> None:
    check(load().cdp_profile_enable(int(enable)), "cdp_profile_enable")


def profile_read() -> dict:
    """{kernel name: (total_ms, launches)} since the last profile_enable()."""
    lib = load()
    out = {}
    for name, kid in KERNEL_IDS.items():
        ms, n = ctypes.c_double(0.0), c_int32(0)
        check(lib.cdp_profile_read(kid, ctypes.byref(ms), ctypes.byref(n)), "cdp_profile_read")
        out[name] = (ms.value, n.value)
    return out


__all__ = ["PhotoArgs", "PhotoHeads", "SIGNATURES", "load", "check", "library_path", "NativeError",
           "MAX_LEVELS", "MAX_BATCH_PER_LAUNCH", "MAX_FLOW_MAPS", "c_uint8"]
