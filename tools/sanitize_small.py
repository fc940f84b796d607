"""A few small launches of every kernel family, compared with the oracle; small enough to run under compute-sanitizer
(racecheck / memcheck / synccheck) where the pool allows it (the round-2 GPU pool does not):
    [compute-sanitizer --tool racecheck] python -m tools.sanitize_small
Shapes: 16-byte aligned rows (aligned kernels), odd pitches and an unaligned base pointer (any-alignment variants),
several column strips and row bands each; results are compared with the oracle as well."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device
from oracle import oracle as O

rng = np.random.default_rng(5)
bad = 0
for (h, w, c, shift) in ((70, 768, 3, 0), (70, 803, 3, 1), (53, 1101, 1, 3), (40, 512, 4, 0), (45, 300, 4, 2)):
    img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    buf = torch.zeros(img.size + 64, dtype=torch.uint8, device="cuda")
    x = buf[16 + shift: 16 + shift + img.size].view(h, w, c)
    x.copy_(torch.from_numpy(img))
    cases = [("sobel1", lambda: device.sobel_edge_detection(x, 1), lambda: O.sobel(img, 1)),
             ("sobel2", lambda: device.sobel_edge_detection(x, 2), lambda: O.sobel(img, 2)),
             ("box3", lambda: device.box_blur(x, 3, 2), lambda: O.box_blur(img, 3)),
             ("box12", lambda: device.box_blur(x, 12, 1), lambda: O.box_blur(img, 12)),
             ("gauss3", lambda: device.gaussian_blur(x, 2.0, 3, 1), lambda: O.gaussian_blur(img, 2.0, 3)),
             ("gauss1", lambda: device.gaussian_blur(x, 0.8, 1, 1), lambda: O.gaussian_blur(img, 0.8, 1)),
             ("gauss9", lambda: device.gaussian_blur(x, 3.0, 9, 1), lambda: O.gaussian_blur(img, 3.0, 9)),
             ("gauss20", lambda: device.gaussian_blur(x, 6.0, 20, 1), lambda: O.gaussian_blur(img, 6.0, 20))]
    for name, run, want in cases:
        got = run().cpu().numpy()
        ok = np.array_equal(got, want())
        bad += not ok
        print(f"{h}x{w}x{c} shift {shift} {name}: {'ok' if ok else 'MISMATCH'}", flush=True)
print("mismatches", bad)
