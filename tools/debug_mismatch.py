"""Where do GPU and oracle differ?  python -m tools.debug_mismatch gaussian H W C R [shift_out]"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpu_image_processing_b200 import device
from oracle import oracle as O

kind, h, w, c, r = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
so = int(sys.argv[6]) if len(sys.argv) > 6 else 0
rng = np.random.default_rng(5)
img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
x = torch.from_numpy(img).cuda()
buf = torch.zeros(img.size + 64, dtype=torch.uint8, device="cuda")
y = buf[so:so + img.size].view(h, w, c)
if kind == "gaussian":
    device.gaussian_blur(x, 2.0, r, out=y); want = O.gaussian_blur(img, 2.0, r)
elif kind == "box":
    device.box_blur(x, r, out=y); want = O.box_blur(img, r)
else:
    device.sobel_edge_detection(x, r, out=y); want = O.sobel(img, r)
torch.cuda.synchronize()
got = y.cpu().numpy()
bad = np.argwhere(got.reshape(h, w * c) != want.reshape(h, w * c))
print("bad bytes", len(bad), "of", img.size, "pitch", w * c)
if len(bad):
    cols = sorted(set(int(b[1]) for b in bad)); rows = sorted(set(int(b[0]) for b in bad))
    print("cols", cols[:60], "... n", len(cols))
    print("rows", rows[:40], "... n", len(rows))
    for b in bad[:10]:
        print(tuple(int(v) for v in b), "got", got.reshape(h, -1)[b[0], b[1]], "want", want.reshape(h, -1)[b[0], b[1]])
