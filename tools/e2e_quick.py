"""The bench's end-to-end step alone: 512 1080p RGB frames through the three gip_*_host calls, pinned buffers.
    python -m tools.e2e_quick [frames] [reps]"""
import ctypes, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import _lib
L = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
FW, FH, FC = 1920, 1080, 3
torch.cuda.init()
hx = torch.randint(0, 256, (n, FH, FW, FC), dtype=torch.uint8).pin_memory()
hy = torch.empty_like(hx).pin_memory()
m = _lib.Metrics()
def step():
    t = []
    for call in (lambda: L.gip_gaussian_blur_host(hx.data_ptr(), hy.data_ptr(), FW, FH, FC, n, 2.0, 3, 3, ctypes.byref(m)),
                 lambda: L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), FW, FH, FC, n, 3, 2, ctypes.byref(m)),
                 lambda: L.gip_sobel_host(hx.data_ptr(), hy.data_ptr(), FW, FH, FC, n, 1, ctypes.byref(m))):
        t0 = time.perf_counter(); _lib.check(call()); t.append((time.perf_counter() - t0) * 1e3)
    return t
step()
for _ in range(reps):
    t = step()
    tot = sum(t)
    print(f"{n} frames: gaussian {t[0]:.1f} box {t[1]:.1f} sobel {t[2]:.1f} ms  -> {3 * n * FW * FH / tot / 1e3:.0f} Mpix/s, {n * FW * FH * FC / (tot / 3) / 1e6:.1f} GB/s each way", flush=True)
