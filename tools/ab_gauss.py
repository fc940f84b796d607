"""A/B helper: Gaussian r = 3 timing on a c4 slice and the c3 shape, plus an output checksum (compare across builds).
    python -m tools.ab_gauss [frames]"""
import os, sys, hashlib
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator(device="cuda").manual_seed(2)
def run(x, reps, sigma=2.0, r=3):
    y = torch.empty_like(x)
    for _ in range(2):
        device.gaussian_blur(x, sigma, r, 1, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        device.gaussian_blur(x, sigma, r, 1, out=y)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, y
x = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
ms, y = run(x, 5)
h = hashlib.sha1(y[:8].cpu().numpy().tobytes()).hexdigest()[:12]
print(f"c4 {n} frames r=3: {ms:7.3f} ms -> {ms * 4096 / n:6.2f} ms per 4096 frames  sha {h}", flush=True)
del x, y
xs = [torch.randint(0, 256, (4320, 7680, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(8)]
for r, s in ((3, 2.0), (1, 0.8), (4, 2.5)):
    tot = 0.0
    for xi in xs:
        ms, y = run(xi, 3, s, r); tot += ms
    h = hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"c3 shape r={r}: {tot / len(xs) * 1000:7.1f} us  sha {h}", flush=True)
