"""gpu_image_processing_b200 -- B200 (sm_100a) implementation of the per-pixel filter hot path of
Pfactorial01/gpu_image_processing: Gaussian blur, box blur and Sobel on u8 gray/RGB/RGBA.

    gpu_filters        the reference's Python module (numpy in, dict out)
    device             torch CUDA tensors in, tensors out (batched, stream-ordered)
    bands              row-band / batch partitioning across GPUs (torch.distributed plumbing)
    _lib               ctypes view of the C ABI, include/gip_b200.h
"""
from . import _lib  # noqa: F401
from . import gpu_filters  # noqa: F401

__all__ = ["_lib", "gpu_filters"]
