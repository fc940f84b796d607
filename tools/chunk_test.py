"""Sobel over the 4096-frame c4 stream as one launch or as chunks of the same resident tensor (the working set, not the
launch size, decides whether the L2 prefetch pays: DESIGN.md 4.2).  python tools/chunk_test.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_image_processing_b200 import device
n = 4096
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
y = torch.empty_like(x)
def run(chunk):
    for s in range(0, n, chunk):
        device.sobel_edge_detection(x[s:s + chunk], 1, out=y[s:s + chunk])
for chunk in (4096, 2048, 1024, 512, 256):
    run(chunk); run(chunk); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(4): run(chunk)
    e1.record(); torch.cuda.synchronize()
    print(f"chunk {chunk}: {e0.elapsed_time(e1) / 4:.3f} ms per 4096 frames", flush=True)
