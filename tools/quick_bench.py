"""Quick device-resident timing of the filters (CUDA events, rotating buffers > L2).
    python -m tools.quick_bench [box|gaussian|sobel|all] [--radii 1,3,7,15,31]"""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device, _lib


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(i); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    # back to back: `reps` launches between one pair of events (what bench.py times)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return ts[len(ts) // 2], e0.elapsed_time(e1) / reps


def run(kind, shape, nbuf, radii):
    h, w, c = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randint(0, 256, (h, w, c), dtype=torch.uint8, device="cuda", generator=g) for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    nbytes = h * w * c
    for r in radii:
        if kind == "box":
            fn = lambda i: device.box_blur(xs[i % nbuf], r, out=ys[i % nbuf])
        elif kind == "gaussian":
            fn = lambda i: device.gaussian_blur(xs[i % nbuf], max(r / 3.0, 0.5), r, out=ys[i % nbuf])
        else:
            fn = lambda i: device.sobel_edge_detection(xs[i % nbuf], r, out=ys[i % nbuf])
        med, b2b = timeit(fn)
        print(f"{kind:8s} {h}x{w}x{c} r/level={r:2d}: single {med*1e3:8.1f} us  back-to-back {b2b*1e3:8.1f} us  "
              f"{2*nbytes/b2b/1e6:7.1f} GB/s (alg)  {h*w/b2b/1e3:9.1f} Mpix/s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    radii = [1, 3, 7, 15, 31]
    for a in sys.argv:
        if a.startswith("--radii="):
            radii = [int(v) for v in a.split("=")[1].split(",")]
    path = 0
    for a in sys.argv:
        if a.startswith("--path="):
            path = int(a.split("=")[1])
    _lib.load().gip_set_path(path)
    print(torch.cuda.get_device_name(0), _lib.load().gip_version().decode(), "path", path)
    if what in ("box", "all"):
        run("box", (4096, 4096, 4), 4, radii)
        run("box", (4320, 7680, 3), 3, radii)
        run("box", (4096, 4096, 1), 8, radii)
    if what == "small":                     # latency of small single images
        for shape in ((1080, 1920, 3), (2146, 3239, 3), (256, 4096, 4), (480, 640, 3)):
            run("box", shape, 16, radii)
            run("sobel", shape, 16, [1])
            run("gaussian", shape, 16, [3])
    if what == "gsweep":                    # Gaussian over the radii on the 8K RGB shape
        run("gaussian", (4320, 7680, 3), 3, radii)
    if what in ("gaussian", "all"):
        run("gaussian", (2146, 3239, 3), 8, [3])
        run("gaussian", (4320, 7680, 3), 3, [3, 15])
    if what in ("sobel", "all"):
        run("sobel", (4320, 7680, 3), 3, [1, 2])
        run("sobel", (4096, 4096, 4), 4, [1, 2])
        run("sobel", (4096, 4096, 1), 8, [1])
