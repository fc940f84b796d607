"""Does the distance between the input and the output buffer matter?  Sobel / box / Gaussian over the 4096-frame c4
stream with the output placed `off` bytes past where torch would put it.  python tools/alias_test.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_image_processing_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda", generator=g)
ybase = torch.empty(x.numel() + (64 << 20), dtype=torch.uint8, device="cuda")
print("in", hex(x.data_ptr()), "out base", hex(ybase.data_ptr()), "distance", hex(ybase.data_ptr() - x.data_ptr()))
for off in (0, 4096, 1 << 16, (1 << 20) + 4096, (3 << 20) + (1 << 16), (17 << 20) + 8192 * 3):
    y = ybase[off:off + x.numel()].view(x.shape)
    res = []
    for name, fn in (("sobel", lambda: device.sobel_edge_detection(x, 1, out=y)), ("box", lambda: device.box_blur(x, 3, 1, out=y)),
                     ("gauss", lambda: device.gaussian_blur(x, 2.0, 3, 1, out=y))):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(4): fn()
        e1.record(); torch.cuda.synchronize()
        res.append(f"{name} {e0.elapsed_time(e1) / 4:.3f} ms")
    print(f"out offset {off:>10d}: " + "  ".join(res), flush=True)
