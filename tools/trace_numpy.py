"""GIP_VERBOSE=1 python -m tools.trace_numpy : per-chunk timeline of one gpu_filters.box_blur(ndarray) call (64 MiB image)."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import gpu_filters
img = np.random.default_rng(0).integers(0, 256, (4096, 4096, 4), dtype=np.uint8)
for i in range(3):
    r = gpu_filters.box_blur(img, radius=3, level=2)
sys.stderr.write("==== timed call\n")
t0 = time.perf_counter(); r = gpu_filters.box_blur(img, radius=4, level=2); t1 = time.perf_counter()
sys.stderr.write("wall %.3f ms\n" % ((t1 - t0) * 1e3))
