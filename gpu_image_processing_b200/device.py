"""Device-resident API: torch CUDA uint8 tensors in, tensors out, stream-ordered.

    gaussian_blur(x, sigma=2.0, radius=3, level=1, out=None)
    box_blur(x, radius=3, level=1, out=None)
    sobel_edge_detection(x, level=1, out=None)

`x` is (H, W, C) or a batch (N, H, W, C), uint8, contiguous, on a CUDA device; C in {1,3,4}.
The launch goes to torch's current stream of x's device through gip_*_async (include/gip_b200.h);
nothing is copied and nothing synchronises.  Parameters mean what they mean in the reference's
`gpu_filters` module (bindings.cpp:243-277).  torch is only the owner of the memory and the
stream here; the arithmetic is in libgip_b200.so.
"""
from __future__ import annotations

import torch

from . import _lib

NAIVE, SHARED_MEMORY, TEXTURE_MEMORY = 1, 2, 3


def _shape(x: torch.Tensor):
    if not x.is_cuda:
        raise RuntimeError("device API needs a CUDA tensor (there is no CPU fallback)")
    if x.dtype != torch.uint8:
        raise RuntimeError("image tensor must be uint8")
    if x.dim() == 3:
        n, (h, w, c) = 1, x.shape
    elif x.dim() == 4:
        n, h, w, c = x.shape
    else:
        raise RuntimeError("Input must be (H, W, C) or (N, H, W, C)")
    if c not in (1, 3, 4):
        raise RuntimeError("Channels must be 1, 3, or 4")
    if not x.is_contiguous():
        raise RuntimeError("image tensor must be contiguous")
    return n, h, w, c


def _out(x, out):
    if out is None:
        return torch.empty_like(x)
    if out.shape != x.shape or out.dtype != torch.uint8 or out.device != x.device or not out.is_contiguous():
        raise RuntimeError("out must match the input's shape, dtype, device and be contiguous")
    return out


def _stream(x):
    return torch.cuda.current_stream(x.device).cuda_stream


def gaussian_blur(x, sigma: float = 2.0, radius: int = 3, level: int = 1, out=None):
    n, h, w, c = _shape(x)
    if level not in (1, 2):
        raise RuntimeError("Level must be 1 (naive) or 2 (texture_memory) for Gaussian blur")
    out = _out(x, out)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().gip_gaussian_blur_async(
            x.data_ptr(), out.data_ptr(), w, h, c, n, float(sigma), int(radius),
            NAIVE if level == 1 else TEXTURE_MEMORY, _stream(x)))
    return out


def box_blur(x, radius: int = 3, level: int = 1, out=None):
    n, h, w, c = _shape(x)
    if level not in (1, 2):
        raise RuntimeError("Level must be 1 (naive) or 2 (shared_memory)")
    out = _out(x, out)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().gip_box_blur_async(
            x.data_ptr(), out.data_ptr(), w, h, c, n, int(radius), int(level), _stream(x)))
    return out


def sobel_edge_detection(x, level: int = 1, out=None):
    n, h, w, c = _shape(x)
    if level not in (1, 2):
        raise RuntimeError("Level must be 1 (naive) or 2 (shared_memory) for Sobel edge detection")
    out = _out(x, out)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().gip_sobel_async(
            x.data_ptr(), out.data_ptr(), w, h, c, n, int(level), _stream(x)))
    return out
