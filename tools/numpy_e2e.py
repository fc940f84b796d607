"""The reference's plugin call on numpy arrays (pageable in, new array out) beside the pinned C-ABI call.
    python -m tools.numpy_e2e"""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gpu_image_processing_b200 import _lib, gpu_filters

L = _lib.load()
rng = np.random.default_rng(0)
for shape in ((4096, 4096, 4), (1080, 1920, 3), (2146, 3239, 3)):
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    h, w, c = shape
    for _ in range(3):
        r = gpu_filters.box_blur(img, radius=3, level=2)
    reps = 10
    t0 = time.perf_counter()
    for i in range(reps):
        r = gpu_filters.box_blur(img, radius=3 + i % 3, level=2)
    ms_np = (time.perf_counter() - t0) / reps * 1e3
    keep = []
    t0 = time.perf_counter()
    for i in range(reps):
        keep.append(gpu_filters.box_blur(img, radius=3 + i % 3, level=2)["image"])     # results kept alive: the pool grows
    ms_keep = (time.perf_counter() - t0) / reps * 1e3
    del keep
    hx = torch.from_numpy(img).pin_memory(); hy = torch.empty_like(hx).pin_memory()
    m = _lib.Metrics()
    L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), w, h, c, 1, 3, 2, ctypes.byref(m))
    t0 = time.perf_counter()
    for i in range(reps):
        L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), w, h, c, 1, 3 + i % 3, 2, ctypes.byref(m))
    ms_pin = (time.perf_counter() - t0) / reps * 1e3
    print(f"{shape}: numpy {ms_np:.3f} ms  numpy(results kept) {ms_keep:.3f} ms  pinned {ms_pin:.3f} ms  ratio {ms_np / ms_pin:.2f}", flush=True)
