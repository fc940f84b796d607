"""Does a small kernel run slower when the GPU is otherwise idle (clock ramp-down)?  Times the 256-row box band
kernel back to back and with host sleeps between launches, and samples the SM clock meanwhile."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpu_image_processing_b200 import _lib

L = _lib.load()
x = torch.randint(0, 256, (256, 4096, 4), dtype=torch.uint8, device="cuda")
y = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream


def one():
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    _lib.check(L.gip_box_blur_async(x.data_ptr(), y.data_ptr(), 4096, 256, 4, 1, 3, 2, st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


def clock():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate", "--format=csv,noheader"],
                          capture_output=True, text=True).stdout.strip()


for _ in range(20):
    one()
busy = sorted(one() for _ in range(200))
print("back to back: median %.1f us" % busy[100], clock())
for gap in (0.0005, 0.002, 0.01, 0.05):
    ts = []
    for _ in range(40):
        time.sleep(gap)
        ts.append(one())
    ts.sort()
    print("sleep %.1f ms between launches: median %.1f us  min %.1f" % (gap * 1e3, ts[20], ts[0]), clock())
