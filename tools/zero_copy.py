"""Experiment: run the fused kernels straight on pinned host memory (UVA zero-copy) instead of staging through
device buffers -- PCIe read + compute + PCIe write inside one kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpu_image_processing_b200 import _lib

L = _lib.load()
H = W = 4096
C = 4
hx = torch.randint(0, 256, (H, W, C), dtype=torch.uint8).pin_memory()
hy = torch.empty_like(hx).pin_memory()
dx = hx.cuda()
dy = torch.empty_like(dx)
ref = torch.empty_like(dx)
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 3)


out = {}
for r in (3, 16, 31):
    _lib.check(L.gip_box_blur_async(dx.data_ptr(), ref.data_ptr(), W, H, C, 1, r, 2, st))
    for name, a, b in (("dev->dev", dx, dy), ("host->dev", hx, dy), ("dev->host", dx, hy), ("host->host", hx, hy)):
        ms = timed(lambda: _lib.check(L.gip_box_blur_async(a.data_ptr(), b.data_ptr(), W, H, C, 1, r, 2, st)))
        ok = bool((b.cuda() == ref).all()) if b is hy else bool((b == ref).all())
        out[f"box r={r} {name}"] = [ms, ok]
for name, a, b in (("dev->dev", dx, dy), ("host->dev", hx, dy), ("dev->host", dx, hy), ("host->host", hx, hy)):
    out[f"sobel {name}"] = timed(lambda: _lib.check(L.gip_sobel_async(a.data_ptr(), b.data_ptr(), W, H, C, 1, 1, st)))
    out[f"gauss r3 {name}"] = timed(lambda: _lib.check(L.gip_gaussian_blur_async(a.data_ptr(), b.data_ptr(), W, H, C, 1, 2.0, 3, 1, st)))
print(json.dumps(out, indent=0))
