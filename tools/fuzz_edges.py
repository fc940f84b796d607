"""Over-read / over-write hunt: images are placed FLUSH against the start or the end of their own cudaMalloc
allocation (whole 2 MiB pages), so a kernel that reads or writes even a few bytes outside the image is likely to hit
an unmapped page and fault instead of silently reading a neighbour.  Results are still compared with the oracle.
python tools/fuzz_edges.py [seconds] [seed]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuda.bindings import runtime as rt

from gpu_image_processing_b200 import _lib
from oracle import oracle as O

PAGE = 2 << 20


def ck(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA error {int(err)}")
    return res[1] if len(res) > 1 else None


budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
L = _lib.load()
t_end = time.time() + budget
n, fails = 0, []
while time.time() < t_end and len(fails) < 5:
    c = int(rng.choice([1, 3, 4]))
    h, w = int(rng.integers(1, 500)), int(rng.integers(1, 3000))
    kind = str(rng.choice(["box", "gaussian", "sobel"]))
    r = int(rng.choice([1, 2, 3, 5, 8, 15, 16, 17, 31, 33]))
    level = int(rng.choice([1, 2]))
    path = int(rng.integers(0, 8) == 0)
    at_end_in, at_end_out = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    nb = img.size
    size = (nb + PAGE - 1) // PAGE * PAGE
    d_in_base, d_out_base = ck(rt.cudaMalloc(size)), ck(rt.cudaMalloc(size))
    d_in = int(d_in_base) + (size - nb if at_end_in else 0)
    d_out = int(d_out_base) + (size - nb if at_end_out else 0)
    if os.environ.get("FUZZ_TRACE"):
        with open(os.environ["FUZZ_TRACE"], "w") as f:
            f.write(repr(dict(n=n, kind=kind, h=h, w=w, c=c, r=r, level=level, path=path, at_end_in=at_end_in, at_end_out=at_end_out)) + "\n")
    ck(rt.cudaMemcpy(d_in, img.ctypes.data, nb, rt.cudaMemcpyKind.cudaMemcpyHostToDevice))
    L.gip_set_path(path)
    try:
        if kind == "box":
            rc = L.gip_box_blur_async(d_in, d_out, w, h, c, 1, r, level, None); want = O.box_blur(img, r)
        elif kind == "gaussian":
            rc = L.gip_gaussian_blur_async(d_in, d_out, w, h, c, 1, 2.5, r, 1 if level == 1 else 3, None); want = O.gaussian_blur(img, 2.5, r)
        else:
            rc = L.gip_sobel_async(d_in, d_out, w, h, c, 1, level, None); want = O.sobel(img, level)
    finally:
        L.gip_set_path(0)
    got = np.empty_like(img)
    e = rt.cudaMemcpy(got.ctypes.data, d_out, nb, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    n += 1
    if rc != 0 or int(e[0]) != 0 or not np.array_equal(got, want):
        fails.append(dict(kind=kind, rc=int(rc), copy_err=int(e[0]), h=h, w=w, c=c, r=r, level=level, path=path, at_end_in=at_end_in, at_end_out=at_end_out))
        if int(e[0]) != 0:
            break          # the context is gone after a device fault
    ck(rt.cudaFree(d_in_base)); ck(rt.cudaFree(d_out_base))
print({"cases": n, "failures": fails, "seed": seed})
sys.exit(1 if fails else 0)
