"""A few launches of each filter on a 64-frame slice of the c4 stream (1920x1080 RGB), the command profiled under ncu
for the per-frame DRAM traffic of the headline kernels:  python -m tools.prof_c4 [frames]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2)]
ys = [torch.empty_like(x) for x in xs]
for i in range(3):
    device.gaussian_blur(xs[i % 2], 2.0, 3, 2, out=ys[i % 2])
    device.box_blur(xs[i % 2], 3, 2, out=ys[i % 2])
    device.sobel_edge_detection(xs[i % 2], 1, out=ys[i % 2])
torch.cuda.synchronize()
print("ok")
