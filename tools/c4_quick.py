"""Device-resident timing of the three filters on a slice of the c4 stream (N 1920x1080 RGB frames in one launch).
    python -m tools.c4_quick [frames]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
y = torch.empty_like(x)
fns = {"gaussian": lambda: device.gaussian_blur(x, 2.0, 3, 1, out=y), "box": lambda: device.box_blur(x, 3, 1, out=y),
       "sobel_l1": lambda: device.sobel_edge_detection(x, 1, out=y), "sobel_l2": lambda: device.sobel_edge_detection(x, 2, out=y)}
for name, fn in fns.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:9s} {n} frames: {ms:7.3f} ms  {2 * x.numel() / ms / 1e6:7.1f} GB/s (alg)  x{4096 // n} = {ms * 4096 / n:6.2f} ms per 4096 frames", flush=True)
