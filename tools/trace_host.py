"""Wall time of gip_box_blur_host on a pinned 4096x4096 RGBA image (GIP_VERBOSE=1 adds the per-chunk timeline)."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpu_image_processing_b200 import _lib
L = _lib.load(); m = _lib.Metrics()
H = W = 4096; C = 4
hx = torch.randint(0, 256, (H, W, C), dtype=torch.uint8).pin_memory(); hy = torch.empty_like(hx).pin_memory()
reps = 3 if os.environ.get("GIP_VERBOSE") == "1" else 20
for r in (3, 31):
    ts = []
    for i in range(reps):
        t0 = time.perf_counter()
        _lib.check(L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), W, H, C, 1, r, 2, ctypes.byref(m)))
        ts.append(time.perf_counter() - t0)
    print(f"r={r} min {min(ts)*1e3:.3f} ms  median {sorted(ts)[len(ts)//2]*1e3:.3f} ms  kernel_ms {m.time_ms:.3f}", {k: v for k, v in os.environ.items() if k.startswith("GIP_")})
