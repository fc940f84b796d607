"""PCIe experiment 3: which element of upload -> kernel -> download slows the copies (see pcie_pipe.py)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpu_image_processing_b200 import _lib

L = _lib.load()
H = W = 4096
C = 4
N = H * W * C
hx = torch.randint(0, 256, (N,), dtype=torch.uint8).pin_memory()
hy = torch.empty_like(hx).pin_memory()
dx = torch.empty(N, dtype=torch.uint8, device="cuda")
dy = torch.empty_like(dx)
ux = torch.empty(N, dtype=torch.uint8, device="cuda")
uy = torch.empty_like(ux)
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
CYC_PER_US = 1965


def call(variant, chunks=16):
    step = N // chunks
    rows = H // chunks
    ups = []
    if variant == "B":
        with torch.cuda.stream(s2):
            torch.cuda._sleep(250 * CYC_PER_US)
    kes = []
    for k in range(chunks):
        a, b = k * step, (k + 1) * step
        with torch.cuda.stream(s1):
            dx[a:b].copy_(hx[a:b], non_blocking=True)
            e = torch.cuda.Event(); e.record(s1); ups.append(e)

        def compute(j):
            a, b = j * step, (j + 1) * step
            with torch.cuda.stream(s3):
                if variant in "CDEF":
                    s3.wait_event(ups[min(j + 1, chunks - 1)])
                if variant == "C":
                    torch.cuda._sleep(46 * CYC_PER_US)
                elif variant in "DF":
                    _lib.check(L.gip_box_blur_async(dx.data_ptr() + a, dy.data_ptr() + a, W, rows, C, 1, 3, 2, s3.cuda_stream))
                elif variant == "E":
                    _lib.check(L.gip_box_blur_async(ux.data_ptr() + a, uy.data_ptr() + a, W, rows, C, 1, 3, 2, s3.cuda_stream))
                ke = torch.cuda.Event(); ke.record(s3); kes.append(ke)
            with torch.cuda.stream(s2):
                if variant in "CEF":
                    s2.wait_event(ke)
                elif variant == "D" and j >= 1:
                    s2.wait_event(kes[j - 1])
                hy[a:b].copy_(dy[a:b], non_blocking=True)
        if k >= 1:
            compute(k - 1)
    compute(chunks - 1)
    torch.cuda.synchronize()


def wall(fn, reps=20):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return round(min(ts) * 1e3, 3), round(sorted(ts)[len(ts) // 2] * 1e3, 3)


out = {}
for v, what in (("A", "no deps, no kernel"), ("B", "downloads start 0.25 ms late, no deps"), ("C", "deps + 46 us sleep kernel"),
                ("D", "real kernel, download k waits kernel k-1"), ("E", "deps + real kernel on unrelated memory"),
                ("F", "deps + real kernel (the pipeline)")):
    out[v + ": " + what] = wall(lambda: call(v))
print(json.dumps(out, indent=0))
