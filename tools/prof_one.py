"""Run one filter a few times (for ncu): python -m tools.prof_one box 4096 4096 4 3 [reps]"""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device

kind, h, w, c, r = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
g = torch.Generator(device="cuda").manual_seed(1)
xs = [torch.randint(0, 256, (h, w, c), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3)]
ys = [torch.empty_like(x) for x in xs]
for i in range(reps):
    if kind == "box":
        device.box_blur(xs[i % 3], r, out=ys[i % 3])
    elif kind == "gaussian":
        device.gaussian_blur(xs[i % 3], max(r / 3.0, 0.5), r, out=ys[i % 3])
    else:
        device.sobel_edge_detection(xs[i % 3], r, out=ys[i % 3])
torch.cuda.synchronize()
print("ok")
