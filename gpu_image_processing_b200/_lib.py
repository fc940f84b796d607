"""ctypes binding of libgip_b200.so (include/gip_b200.h).  There is no CPU fallback: if the
library is missing this module raises, and every entry point returns the CUDA error of the
device it ran on."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgip_b200.so")

u8p = ctypes.c_void_p
i32, i64, f32 = ctypes.c_int, ctypes.c_int64, ctypes.c_float


class Metrics(ctypes.Structure):
    """PerformanceMetrics, cuda_lib/include/image_filters.h:17-21."""
    _fields_ = [("time_ms", f32), ("bandwidth_gbps", f32), ("fps", f32)]


mp = ctypes.POINTER(Metrics)

# name -> (restype, argtypes); the complete export list of include/gip_b200.h
SIGNATURES = {
    "gip_gaussian_blur": (i32, [u8p, u8p, i32, i32, i32, f32, i32, i32, mp]),
    "gip_box_blur": (i32, [u8p, u8p, i32, i32, i32, i32, i32, mp]),
    "gip_sobel": (i32, [u8p, u8p, i32, i32, i32, i32, mp]),
    "gip_gaussian_blur_async": (i32, [u8p, u8p, i64, i64, i32, i64, f32, i32, i32, ctypes.c_void_p]),
    "gip_box_blur_async": (i32, [u8p, u8p, i64, i64, i32, i64, i32, i32, ctypes.c_void_p]),
    "gip_sobel_async": (i32, [u8p, u8p, i64, i64, i32, i64, i32, ctypes.c_void_p]),
    "gip_gaussian_blur_band": (i32, [u8p, u8p, u8p, u8p, i64, i64, i32, i64, i64, i64, i64, f32, i32, i32, ctypes.c_void_p]),
    "gip_box_blur_band": (i32, [u8p, u8p, u8p, u8p, i64, i64, i32, i64, i64, i64, i64, i32, i32, ctypes.c_void_p]),
    "gip_sobel_band": (i32, [u8p, u8p, u8p, u8p, i64, i64, i32, i64, i64, i64, i64, i32, ctypes.c_void_p]),
    "gip_gaussian_blur_host": (i32, [u8p, u8p, i64, i64, i32, i64, f32, i32, i32, mp]),
    "gip_box_blur_host": (i32, [u8p, u8p, i64, i64, i32, i64, i32, i32, mp]),
    "gip_sobel_host": (i32, [u8p, u8p, i64, i64, i32, i64, i32, mp]),
    "gip_host_alloc": (i32, [i64, ctypes.POINTER(ctypes.c_void_p)]),
    "gip_host_free": (i32, [ctypes.c_void_p]),
    "gip_device_alloc": (i32, [i64, ctypes.POINTER(ctypes.c_void_p)]),
    "gip_device_free": (i32, [ctypes.c_void_p]),
    "gip_ipc_export": (i32, [ctypes.c_void_p, ctypes.c_void_p]),
    "gip_ipc_open": (i32, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "gip_ipc_close": (i32, [ctypes.c_void_p]),
    "gip_enable_peer_access": (i32, [i32]),
    "gip_gaussian_weights": (i32, [ctypes.POINTER(f32), i32, f32]),
    "gip_error_string": (ctypes.c_char_p, [i32]),
    "gip_launch_count": (i64, []),
    "gip_release_cache": (i32, []),
    "gip_version": (ctypes.c_char_p, []),
    "gip_set_path": (i32, [i32]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m gpu_image_processing_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class CudaError(RuntimeError):
    """Same text as the reference binding's error (bindings.cpp:68): "CUDA error: <string>"."""

    def __init__(self, code: int):
        self.code = int(code)
        msg = load().gip_error_string(self.code)
        super().__init__("CUDA error: " + (msg.decode() if msg else str(code)))


def check(code: int) -> None:
    if code != 0:
        raise CudaError(code)
