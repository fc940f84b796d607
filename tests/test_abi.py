"""CPU tests of the boundary: the shared library loads, exports every symbol that
include/gip_b200.h declares plus the reference's three mangled C++ entry points, and the host
logic (argument checking, error text, no CPU fallback) behaves like the reference binding."""
import ctypes
import os
import re

import numpy as np
import pytest

from gpu_image_processing_b200 import _lib, gpu_filters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    header = open(os.path.join(ROOT, "include", "gip_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gip_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_reference_cxx_symbols_are_exported():
    """Mangled names of cuda_lib/include/image_filters.h:46-112."""
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in ("_Z12gaussianBlurPhS_iiifi17OptimizationLevelP18PerformanceMetrics",
                "_Z7boxBlurPhS_iiii17OptimizationLevelP18PerformanceMetrics",
                "_Z18sobelEdgeDetectionPhS_iii17OptimizationLevelP18PerformanceMetrics"):
        assert getattr(L, sym) is not None


def test_weights_entry_point_matches_oracle():
    from oracle import oracle as O
    for r, s in [(0, 1.0), (3, 2.0), (15, 5.0), (31, 9.5)]:
        w = (ctypes.c_float * (2 * r + 1))()
        assert _lib.load().gip_gaussian_weights(w, r, s) == 0
        assert np.array_equal(np.frombuffer(w, dtype=np.float32), O.gaussian_weights(r, s))
    assert _lib.load().gip_gaussian_weights(None, 3, 2.0) != 0
    assert _lib.load().gip_gaussian_weights((ctypes.c_float * 7)(), 3, -1.0) != 0


def test_module_surface_matches_reference_binding():
    assert (gpu_filters.NAIVE, gpu_filters.SHARED_MEMORY, gpu_filters.TEXTURE_MEMORY) == (1, 2, 3)
    import inspect
    assert list(inspect.signature(gpu_filters.gaussian_blur).parameters) == ["image", "sigma", "radius", "level"]
    assert list(inspect.signature(gpu_filters.box_blur).parameters) == ["image", "radius", "level"]
    assert list(inspect.signature(gpu_filters.sobel_edge_detection).parameters) == ["image", "level"]
    d = {k: v.default for k, v in inspect.signature(gpu_filters.gaussian_blur).parameters.items()}
    assert d["sigma"] == 2.0 and d["radius"] == 3 and d["level"] == 1


def test_argument_errors_match_reference_text():
    img = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(RuntimeError, match="Input must be 3D array"):
        gpu_filters.gaussian_blur(np.zeros((4, 4), np.uint8))
    with pytest.raises(RuntimeError, match="Channels must be 1, 3, or 4"):
        gpu_filters.box_blur(np.zeros((4, 4, 2), np.uint8))
    with pytest.raises(RuntimeError, match=r"Level must be 1 \(naive\) or 2 \(texture_memory\) for Gaussian blur"):
        gpu_filters.gaussian_blur(img, level=3)
    with pytest.raises(RuntimeError, match=r"Level must be 1 \(naive\) or 2 \(shared_memory\)"):
        gpu_filters.box_blur(img, level=0)
    with pytest.raises(RuntimeError, match="for Sobel edge detection"):
        gpu_filters.sobel_edge_detection(img, level=5)


def test_level_and_argument_validation_in_the_c_layer():
    L = _lib.load()
    m = _lib.Metrics()
    # unsupported level -> cudaErrorNotSupported (801) before anything touches the device
    assert L.gip_gaussian_blur(1, 1, 8, 8, 3, 2.0, 3, 2, ctypes.byref(m)) == 801
    assert L.gip_box_blur(1, 1, 8, 8, 3, 3, 3, ctypes.byref(m)) == 801
    assert L.gip_sobel(1, 1, 8, 8, 3, 4, ctypes.byref(m)) == 801
    assert b"not supported" in L.gip_error_string(801)


@pytest.mark.skipif(_has_cuda(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    """On a machine without a CUDA device the product path must fail loudly, not compute on the CPU."""
    img = np.zeros((8, 8, 3), np.uint8)
    for call in (lambda: gpu_filters.gaussian_blur(img), lambda: gpu_filters.box_blur(img),
                 lambda: gpu_filters.sobel_edge_detection(img)):
        with pytest.raises(RuntimeError, match="CUDA error"):
            call()


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gpu_image_processing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"liboracle", r"gipo_", r"oracle[/.]_ref", r"filters_oracle"):
                    assert not re.search(pat, text), (pat, os.path.join(dirpath, f))
