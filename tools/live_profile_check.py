"""Run the ncu side-car (profiling/ncu_profiler.py) end to end on the GPU box: one plain filter call, then the
same filter under ncu through profile_kernel_with_ncu(), and print the shape of what comes back."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gpu_image_processing_b200 import gpu_filters
from gpu_image_processing_b200.profiling import ncu_profiler as P

img = np.random.default_rng(0).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
r = gpu_filters.gaussian_blur(img, sigma=2.0, radius=3, level=2)
print("plain call ok, time_ms", r["time_ms"])
print("ncu available", P.check_ncu_available())
res = P.profile_kernel_with_ncu(img, "gaussian", 2, sigma=2.0, radius=3)
short = {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in list(v.items())[:4]}) for k, v in res.items()}
print(json.dumps(short, default=str)[:3000])
common = P.get_common_ncu_metrics(res.get("metrics", res), ncu_data=res.get("metrics", res)) if isinstance(res, dict) else {}
print("common", json.dumps(common, default=str)[:800])
