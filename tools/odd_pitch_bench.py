"""c1-shape (3239 x 2146 RGB, odd pitch) against the nearest 16-byte aligned shape (3248 x 2146), all three filters.
    python -m tools.odd_pitch_bench"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device


def t(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


g = torch.Generator(device="cuda").manual_seed(3)
for w in (3239, 3240, 3248):
    xs = [torch.randint(0, 256, (2146, w, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(8)]
    ys = [torch.empty_like(x) for x in xs]
    res = {"sobel_l1": t(lambda i: device.sobel_edge_detection(xs[i % 8], 1, out=ys[i % 8])),
           "sobel_l2": t(lambda i: device.sobel_edge_detection(xs[i % 8], 2, out=ys[i % 8])),
           "box_r3": t(lambda i: device.box_blur(xs[i % 8], 3, 2, out=ys[i % 8])),
           "box_r12": t(lambda i: device.box_blur(xs[i % 8], 12, 2, out=ys[i % 8])),
           "gauss_r3": t(lambda i: device.gaussian_blur(xs[i % 8], 2.0, 3, 1, out=ys[i % 8]))}
    print(w, {k: round(v, 1) for k, v in res.items()}, "us", flush=True)
