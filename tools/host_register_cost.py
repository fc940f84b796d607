"""How long does page-locking a caller's numpy array in place take (cudaHostRegister / cudaHostUnregister), against
copying it into a pinned staging buffer with N threads?   python -m tools.host_register_cost"""
import time, ctypes, threading
import numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init()
for mb in (6, 21, 64, 256):
    n = mb << 20
    a = np.random.default_rng(0).integers(0, 256, size=n, dtype=np.uint8)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(a.ctypes.data, n, 0)
        t1 = time.perf_counter()
        rt.cudaHostUnregister(a.ctypes.data)
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1))
    reg = min(t[0] for t in ts) * 1e3; unreg = min(t[1] for t in ts) * 1e3
    pinned = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    res = {}
    for nt in (1, 4, 8, 16):
        def work(j):
            lo, hi = n * j // nt, n * (j + 1) // nt
            ctypes.memmove(pinned.ctypes.data + lo, a.ctypes.data + lo, hi - lo)
        best = 1e9
        for _ in range(5):
            th = [threading.Thread(target=work, args=(j,)) for j in range(nt)]
            t0 = time.perf_counter()
            for t in th: t.start()
            for t in th: t.join()
            best = min(best, time.perf_counter() - t0)
        res[nt] = round(best * 1e3, 3)
    print(f"{mb} MiB: register {reg:.3f} ms (rc {rc}) unregister {unreg:.3f} ms; memcpy to pinned (ms by threads) {res}", flush=True)
