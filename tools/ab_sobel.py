"""A/B helper: Sobel timing on a c4 slice, the c3 shape and 4096^2 RGBA, plus output checksums (compare across builds).
    python -m tools.ab_sobel [frames]"""
import os, sys, hashlib
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator(device="cuda").manual_seed(2)
def run(x, reps, level):
    y = torch.empty_like(x)
    for _ in range(2):
        device.sobel_edge_detection(x, level, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        device.sobel_edge_detection(x, level, out=y)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, y
x = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
for level in (1, 2):
    ms, y = run(x, 5, level)
    h = hashlib.sha1(y[:8].cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"c4 {n} frames level {level}: {ms:7.3f} ms -> {ms * 4096 / n:6.2f} ms per 4096 frames  sha {h}", flush=True)
del x, y
for shape in ((4320, 7680, 3), (4096, 4096, 4)):
    xs = [torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda", generator=g) for _ in range(8)]
    for level in (1, 2):
        tot = 0.0
        for xi in xs:
            ms, y = run(xi, 3, level); tot += ms
        h = hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:12]
        print(f"{shape} level {level}: {tot / len(xs) * 1000:7.1f} us  sha {h}", flush=True)
