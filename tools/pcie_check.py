"""Where does the host-buffer path's time go?  PCIe copy rates (one direction, both at once) for a
64 MiB pinned image and the gip_box_blur_host call beside them.  Run on the GPU box:
    python tools/pcie_check.py            (GIP_HOST_CHUNK_KB=... to try other pipeline chunk sizes)
"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from gpu_image_processing_b200 import _lib  # noqa: E402


def wall(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    H = W = 4096
    C = 4
    nbytes = H * W * C
    hx = torch.randint(0, 256, (H, W, C), dtype=torch.uint8).pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    dx = torch.empty((H, W, C), dtype=torch.uint8, device="cuda")
    dy = torch.empty_like(dx)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            dx.copy_(hx, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            hy.copy_(dy, non_blocking=True)

    def both():
        h2d()
        d2h()

    out = {"bytes": nbytes}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
        ms = wall(fn)
        out[name] = {"ms": ms, "GB/s_each_way": nbytes / ms / 1e6}

    L = _lib.load()
    m = _lib.Metrics()
    for r in (1, 3, 16, 31):
        ms = wall(lambda: _lib.check(L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), W, H, C, 1, r, 2, ctypes.byref(m))))
        out[f"box_host_r{r}"] = {"ms": ms, "Mpix/s": H * W / ms / 1e3, "kernel_ms": m.time_ms}
    ms = wall(lambda: _lib.check(L.gip_sobel_host(hx.data_ptr(), hy.data_ptr(), W, H, C, 1, 1, ctypes.byref(m))))
    out["sobel_host"] = {"ms": ms, "Mpix/s": H * W / ms / 1e3, "kernel_ms": m.time_ms}
    # pageable caller memory (what numpy arrays from the REST path are)
    px = torch.randint(0, 256, (H, W, C), dtype=torch.uint8)
    py = torch.empty_like(px)
    ms = wall(lambda: _lib.check(L.gip_box_blur_host(px.data_ptr(), py.data_ptr(), W, H, C, 1, 3, 2, ctypes.byref(m))), reps=5)
    out["box_host_r3_pageable"] = {"ms": ms, "Mpix/s": H * W / ms / 1e3}
    out["chunk_kb"] = os.environ.get("GIP_HOST_CHUNK_KB", "default 8192")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
