"""ctypes front-end of the CPU oracle (oracle/filters_oracle.c) and of the reference's own
CUDA library when it was built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

The oracle follows /root/reference/cuda_lib/src/image_filters.cu (see the header of
filters_oracle.c for the line-by-line map).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_image_filters.so")

_u8p = ctypes.POINTER(ctypes.c_uint8)
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when the reference sources are mounted)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "filters_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_float
        L.gipo_gaussian_weights.argtypes = [ctypes.POINTER(f32), i32, f32]
        L.gipo_gaussian_weights.restype = None
        L.gipo_gaussian_blur.argtypes = [_u8p, _u8p, i64, i64, i32, f32, i32, i32]
        L.gipo_box_blur.argtypes = [_u8p, _u8p, i64, i64, i32, i32, i32]
        L.gipo_box_blur_int.argtypes = [_u8p, _u8p, i64, i64, i32, i32, i32]
        L.gipo_sobel.argtypes = [_u8p, _u8p, i64, i64, i32, i32, i32]
        L.gipo_box_round_float.argtypes = [i32, i32, i32]
        L.gipo_max_threads.argtypes = []
        for f in (L.gipo_gaussian_blur, L.gipo_box_blur, L.gipo_box_blur_int, L.gipo_sobel,
                  L.gipo_box_round_float, L.gipo_max_threads):
            f.restype = i32
        _lib = L
    return _lib


def _prep(image: np.ndarray):
    a = np.ascontiguousarray(image, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3:
        raise ValueError("image must be (H, W, C)")
    return a, np.empty_like(a)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def max_threads() -> int:
    return int(lib().gipo_max_threads())


def gaussian_weights(radius: int, sigma: float) -> np.ndarray:
    k = np.empty(2 * radius + 1, dtype=np.float32)
    lib().gipo_gaussian_weights(k.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), radius, sigma)
    return k


def gaussian_blur(image: np.ndarray, sigma: float = 2.0, radius: int = 3, nthreads: int = 0) -> np.ndarray:
    a, o = _prep(image)
    h, w, c = a.shape
    rc = lib().gipo_gaussian_blur(_p(a), _p(o), w, h, c, sigma, radius, nthreads)
    if rc:
        raise RuntimeError(f"oracle gaussian_blur failed rc={rc}")
    return o


def box_blur(image: np.ndarray, radius: int = 3, nthreads: int = 0, integer: bool = False) -> np.ndarray:
    a, o = _prep(image)
    h, w, c = a.shape
    fn = lib().gipo_box_blur_int if integer else lib().gipo_box_blur
    rc = fn(_p(a), _p(o), w, h, c, radius, nthreads)
    if rc:
        raise RuntimeError(f"oracle box_blur failed rc={rc}")
    return o


def sobel(image: np.ndarray, level: int = 1, nthreads: int = 0) -> np.ndarray:
    a, o = _prep(image)
    h, w, c = a.shape
    rc = lib().gipo_sobel(_p(a), _p(o), w, h, c, level, nthreads)
    if rc:
        raise RuntimeError(f"oracle sobel failed rc={rc}")
    return o


def box_round_float(s: int, k: int, fused: bool) -> int:
    return int(lib().gipo_box_round_float(s, k, 1 if fused else 0))


# ----------------------------------------------------------------------------------------
# The reference's own library (oracle/_ref), callable only where a GPU is present.
# Symbols are the C++-mangled names of cuda_lib/include/image_filters.h:46-112.
# ----------------------------------------------------------------------------------------
class _RefMetrics(ctypes.Structure):
    _fields_ = [("time_ms", ctypes.c_float), ("bandwidth_gbps", ctypes.c_float), ("fps", ctypes.c_float)]


_REF_SYMS = {
    "gaussian": "_Z12gaussianBlurPhS_iiifi17OptimizationLevelP18PerformanceMetrics",
    "box": "_Z7boxBlurPhS_iiii17OptimizationLevelP18PerformanceMetrics",
    "sobel": "_Z18sobelEdgeDetectionPhS_iii17OptimizationLevelP18PerformanceMetrics",
}
_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def ref_lib() -> ctypes.CDLL:
    global _ref
    if _ref is None:
        if not ref_available():
            raise FileNotFoundError(f"{_REF_PATH} missing: run `make -C oracle ref` where /root/reference is mounted")
        R = ctypes.CDLL(_REF_PATH)
        vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        mp = ctypes.POINTER(_RefMetrics)
        g = getattr(R, _REF_SYMS["gaussian"]); g.argtypes = [vp, vp, i32, i32, i32, f32, i32, i32, mp]; g.restype = i32
        b = getattr(R, _REF_SYMS["box"]); b.argtypes = [vp, vp, i32, i32, i32, i32, i32, mp]; b.restype = i32
        s = getattr(R, _REF_SYMS["sobel"]); s.argtypes = [vp, vp, i32, i32, i32, i32, mp]; s.restype = i32
        _ref = (R, g, b, s)
    return _ref[0]


class _Quiet:
    """The reference printf()s on every call (image_filters.cu:42-47,781,920-923): mute fd 1."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)
        return self

    def __exit__(self, *exc):
        try:
            ctypes.CDLL(None).fflush(None)
        finally:
            os.dup2(self._saved, 1)
            os.close(self._saved)
            os.close(self._null)
        return False


def ref_call(kind: str, d_in: int, d_out: int, w: int, h: int, c: int, level_enum: int,
             sigma: float = 2.0, radius: int = 3):
    """Run the reference's own entry point on device pointers; returns (cudaError, time_ms)."""
    ref_lib()
    _, g, b, s = _ref
    m = _RefMetrics()
    with _Quiet():
        if kind == "gaussian":
            rc = g(d_in, d_out, w, h, c, sigma, radius, level_enum, ctypes.byref(m))
        elif kind == "box":
            rc = b(d_in, d_out, w, h, c, radius, level_enum, ctypes.byref(m))
        elif kind == "sobel":
            rc = s(d_in, d_out, w, h, c, level_enum, ctypes.byref(m))
        else:
            raise ValueError(kind)
    return int(rc), float(m.time_ms)
