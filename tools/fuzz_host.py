"""Randomised parity hunt for the host-buffer entry points (chunked three-stream pipeline, staging threads):
random image sizes around the chunk boundaries, batches, radii, pinned and pageable buffers.
python tools/fuzz_host.py [seconds] [seed]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gpu_image_processing_b200 import _lib
from oracle import oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
L = _lib.load()
m = _lib.Metrics()
t_end = time.time() + budget
n, fails = 0, []
while time.time() < t_end and len(fails) < 5:
    c = int(rng.choice([1, 3, 4]))
    kind = str(rng.choice(["box", "gaussian", "sobel"]))
    r = 1 if kind == "sobel" else int(rng.choice([1, 3, 7, 15, 31, 40]))
    level = int(rng.choice([1, 2]))
    batch = int(rng.choice([1, 1, 1, 2, 5, 9]))
    target = int(rng.choice([300_000, 1_500_000, 4_200_000, 8_400_000, 17_000_000, 30_000_000])) // batch
    w = int(rng.integers(16, 5000))
    h = max(1, target // (w * c) + int(rng.integers(-3, 4)))
    img = rng.integers(0, 256, (batch, h, w, c), dtype=np.uint8)
    pin_in, pin_out = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    x = torch.from_numpy(img.copy())
    y = torch.full(img.shape, 0x3C, dtype=torch.uint8)
    if pin_in:
        x = x.pin_memory()
    if pin_out:
        y = y.pin_memory()
    if kind == "box":
        rc = L.gip_box_blur_host(x.data_ptr(), y.data_ptr(), w, h, c, batch, r, level, ctypes.byref(m))
        want = [O.box_blur(img[i], r) for i in range(batch)]
    elif kind == "gaussian":
        rc = L.gip_gaussian_blur_host(x.data_ptr(), y.data_ptr(), w, h, c, batch, 3.0, r, 1 if level == 1 else 3, ctypes.byref(m))
        want = [O.gaussian_blur(img[i], 3.0, r) for i in range(batch)]
    else:
        rc = L.gip_sobel_host(x.data_ptr(), y.data_ptr(), w, h, c, batch, level, ctypes.byref(m))
        want = [O.sobel(img[i], level) for i in range(batch)]
    n += 1
    got = y.numpy()
    if rc != 0 or not all(np.array_equal(got[i], want[i]) for i in range(batch)):
        fails.append(dict(kind=kind, rc=int(rc), h=h, w=w, c=c, r=r, level=level, batch=batch, pin_in=pin_in, pin_out=pin_out))
print({"cases": n, "failures": fails, "seed": seed})
sys.exit(1 if fails else 0)
