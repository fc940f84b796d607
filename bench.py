#!/usr/bin/env python
"""bench.py -- headline benchmark of the filter hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Headline workload = BASELINE.json configs[3] (c4), the largest configuration that fits one GPU and the one the
path shards on: a stream of 4096 synthetic 1920x1080 RGB u8 frames through ALL THREE filters (Gaussian sigma 2
radius 3, box radius 3 -- the API defaults -- and Sobel level 1).  The frames are cut into contiguous ranges, one per
GPU (bands.shard_range, no data-path collective); one *step* = every rank filters its range with one batched launch
per filter.  Total work is fixed, so N GPUs is STRONG scaling.  25.5 GB in + 25.5 GB out are resident in HBM at N = 1
(inputs are >> the 126 MB L2, nothing is cached between launches).

The JSON line carries
  value        Mpix/s = frames x pixels x 3 filters / step time; CUDA events on the launch stream, max over ranks
  filters      per filter Mpix/s, GB/s and HBM fraction from events inside the same timed region; at N = 1 also the
               other BASELINE shapes: c1 (Gaussian, 3239x2146 RGB), c2 (box sweep r = 1..31, 4096^2 RGBA), c3 (Sobel 8K RGB)
  roofline     the dominant kernel of the step (largest share of the step time) against the measured HBM copy bandwidth
               (MEASURED_PEAKS.json); roofline_per_filter has the same block for every filter's kernel, and for the
               Gaussian the FP32-pipe roofline that actually bounds it
  c5           BASELINE configs[4]: one 32768 x 32768 RGB image, Gaussian radius 15: a single launch at N = 1, row bands
               with NVLink P2P halo rows (bands.BandedImage) at N > 1
  e2e          the same three filters through the host-buffer entry points gip_*_host (what gpu_filters.* calls) from
               pinned host memory on a 512-frame slice of the stream: H2D + kernels + D2H inside the timed region
  e2e_numpy    the reference's actual plugin call, gpu_filters.<filter>(ndarray) -> new ndarray (pageable memory)
  cpu_baseline the oracle (C transcription of the reference's level-1 math, OpenMP, all host cores), bounded sample; N = 1
  reference_cuda_same_gpu   the reference's own level-2 kernels (oracle/_ref, unmodified, sm_100a) on this GPU, every config
`--impl reference` times the CPU port on the same workload (the reference has no CPU implementation of this path).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAMES, FH, FW, FC = 4096, 1080, 1920, 3
FRAME_BYTES = FH * FW * FC
FRAME_PIX = FH * FW
G_SIGMA, G_RADIUS, B_RADIUS, S_LEVEL = 2.0, 3, 3, 1
FILTERS = ("gaussian", "box", "sobel")
KERNELS = {"gaussian": "gip_gauss_fused<3,3>", "box": "gip_box_fused<3,true,true,8>", "sobel": "gip_sobel_fused<3,false,8>"}
E2E_FRAMES = 512
C5 = dict(h=32768, w=32768, c=3, radius=15, sigma=5.0)
C2 = dict(h=4096, w=4096, c=4, radii=list(range(1, 32)))
METRIC = "Mpix/s over Gaussian + box + Sobel on a 4096-frame 1920x1080 RGB u8 stream (aggregate over GPUs)"
CONFIG = {"workload": "c4: 4096 frames 1920x1080 RGB u8, all three filters (Gaussian sigma=2 r=3, box r=3, Sobel level 1), "
                      "contiguous frame ranges per GPU, one batched launch per filter and rank",
          "step": "3 launches per rank (one per filter) over the rank's frames",
          "l2_policy": "inputs larger than L2 (25.5 GB stream, every launch reads its frames from HBM)",
          "sharding": "frames sharded per GPU (bands.shard_range), no data-path collective; fixed total work"}


def traffic_table():
    """dram__bytes_read + dram__bytes_write per frame of each kernel, from the committed ncu captures."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Launch nvidia-smi and wait (at most 5 s) for its first sample, so that short timed regions are covered."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t = time.time()
            while not self.rows and time.time() - t < 5.0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 - 0.05 <= t <= (self.t1 or t) + 0.1)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm),
                "window": "nvidia-smi every 50 ms over the device-timed region, the c5 region and the end-to-end region"}


def host_cores():
    try:    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask the scheduler instead)
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_run(steps, warmup, frames=64):
    """The oracle on all host threads over a bounded sample of the workload: `frames` 1080p RGB frames, three filters."""
    import numpy as np
    from oracle import oracle as O
    cores = host_cores()
    rng = np.random.default_rng(1234)
    fr = [rng.integers(0, 256, size=(FH, FW, FC), dtype=np.uint8) for _ in range(4)]
    calls = {"gaussian": lambda f: O.gaussian_blur(f, G_SIGMA, G_RADIUS, nthreads=cores),
             "box": lambda f: O.box_blur(f, B_RADIUS, nthreads=cores),
             "sobel": lambda f: O.sobel(f, S_LEVEL, nthreads=cores)}
    for _ in range(max(0, min(warmup, 1))):
        for k in FILTERS:
            calls[k](fr[0])
    best, per = None, None
    for _ in range(max(1, steps)):
        t_f = {}
        t0 = time.perf_counter()
        for k in FILTERS:
            t1 = time.perf_counter()
            for i in range(frames):
                calls[k](fr[i % 4])
            t_f[k] = time.perf_counter() - t1
        t = time.perf_counter() - t0
        if best is None or t < best:
            best, per = t, t_f
    mpix = 3 * frames * FRAME_PIX / best / 1e6
    return {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
            "per_filter_Mpix/s": {k: frames * FRAME_PIX / v / 1e6 for k, v in per.items()},
            "sample": f"{frames} of the 4096 1920x1080 RGB frames through all three filters, best of {max(1, steps)}, "
                      f"{best:.2f} s per pass; straight C transcription of the reference's level-1 math "
                      "(oracle/filters_oracle.c), OpenMP over all host threads, one frame per call"}, best


def run_reference_arm(args):
    """--impl reference: the CPU port on the host cores (the reference itself has no CPU path)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 3))
    base, pass_s = cpu_port_run(steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": pass_s * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference arm = CPU port of the reference's level-1 math on the host cores; each step is the "
                    "bounded sample named in cpu_baseline.sample"}
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
def reference_cuda_same_gpu(torch):
    """the reference's own kernels (oracle/_ref, unmodified, compiled for sm_100a) on this GPU, every
    config, level 2 wherever its level 2 is correct; its own time_ms (CUDA events around its kernels only)."""
    out = {}
    try:
        from oracle import oracle as O
        if not O.ref_available():
            return {"unavailable": "oracle/_ref not built"}
        g = torch.Generator(device="cuda").manual_seed(99)

        def img(h, w, c):
            x = torch.randint(0, 256, (h, w, c), dtype=torch.uint8, device="cuda", generator=g)
            return x, torch.empty_like(x)

        def call(kind, x, y, lvl, sigma=0.0, radius=1, reps=3):
            h, w, c = x.shape
            best = None
            for _ in range(reps):
                rc, ms = O.ref_call(kind, x.data_ptr(), y.data_ptr(), w, h, c, lvl, sigma, radius)
                if rc != 0:
                    raise RuntimeError(f"reference {kind} returned {rc}")
                best = ms if best is None else min(best, ms)
            return best

        def entry(ms, npix, nbytes, how):
            return {"ms": ms, "Mpix/s": npix / ms / 1e3, "alg_GB/s": 2 * nbytes / ms / 1e6, "how": how}

        x, y = img(2146, 3239, 3)
        out["c1_gaussian_3239x2146_rgb_s2_r3"] = entry(call("gaussian", x, y, 3, 2.0, 3), x.shape[0] * x.shape[1], x.numel(), "gaussianBlur level 2 (TEXTURE_MEMORY)")
        x, y = img(C2["h"], C2["w"], C2["c"])
        tot = sum(call("box", x, y, 2 if r <= 16 else 1, 0.0, r, reps=1) for r in C2["radii"])
        out["c2_box_sweep_r1_31_4096sq_rgba"] = entry(tot, len(C2["radii"]) * C2["h"] * C2["w"], len(C2["radii"]) * x.numel(),
                                                      "boxBlur level 2 for r <= 16, level 1 above (its level 2 is wrong for r > 16)")
        x, y = img(4320, 7680, 3)
        out["c3_sobel_7680x4320_rgb"] = entry(call("sobel", x, y, 2), 4320 * 7680, x.numel(), "sobelEdgeDetection level 2")
        out["c3_shape_gaussian_s2_r3"] = entry(call("gaussian", x, y, 3, 2.0, 3), 4320 * 7680, x.numel(), "gaussianBlur level 2")
        del x, y
        nf = 16
        xs = torch.randint(0, 256, (nf, FH, FW, FC), dtype=torch.uint8, device="cuda", generator=g)
        ys = torch.empty_like(xs)
        c4 = {}
        for kind, lvl, sigma, radius in (("gaussian", 3, G_SIGMA, G_RADIUS), ("box", 2, 0.0, B_RADIUS), ("sobel", 2, 0.0, 1)):
            call(kind, xs[0], ys[0], lvl, sigma, radius, reps=1)
            ms = sum(call(kind, xs[i], ys[i], lvl, sigma, radius, reps=1) for i in range(nf))
            c4[kind] = entry(ms, nf * FRAME_PIX, nf * FRAME_BYTES, f"{nf} frames, one call per frame (the reference has no batch call), level 2, kernel time only")
        c4["all_three_Mpix/s"] = 3 * nf * FRAME_PIX / sum(v["ms"] for v in c4.values()) / 1e3
        out["c4_1920x1080_rgb_frames"] = c4
        del xs, ys
        rows = 4096                                   # a band of c5: the reference's int arithmetic cannot address the whole image
        x, y = img(rows, C5["w"], C5["c"])
        out["c5_band_gaussian_32768x4096_rgb_s5_r15"] = entry(call("gaussian", x, y, 3, C5["sigma"], C5["radius"], reps=2), rows * C5["w"], x.numel(),
                                                               "gaussianBlur level 2 on a 4096-row band (1/8 of the image; 32-bit sizes, image_filters.cu:760)")
        del x, y
        torch.cuda.synchronize()
    except Exception as e:  # the baseline must never take the bench down
        out["error"] = repr(e)
    return out


def other_configs(torch, device_mod, peak):
    """Device-resident timings of the other BASELINE shapes (c1, c2, c3): reported, not the headline."""
    out = {}
    g = torch.Generator(device="cuda").manual_seed(7)

    def timed(fn, nbytes, npix, reps=12):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 2 * nbytes / ms / 1e6
        return {"us": ms * 1e3, "Mpix/s": npix / ms / 1e3, "alg_GB/s": gbs, "frac_of_hbm_peak": gbs / peak}

    for name, (h, w, c), nb, call in (
            ("c1_gaussian_3239x2146_rgb_s2_r3", (2146, 3239, 3), 8, lambda x, y: device_mod.gaussian_blur(x, 2.0, 3, 2, out=y)),
            # the same odd-pitch shape (9717-byte rows) through the other two filters: the any-alignment paths
            ("c1_shape_box_r3_odd_pitch", (2146, 3239, 3), 8, lambda x, y: device_mod.box_blur(x, 3, 2, out=y)),
            ("c1_shape_sobel_odd_pitch", (2146, 3239, 3), 8, lambda x, y: device_mod.sobel_edge_detection(x, 1, out=y)),
            ("c3_sobel_7680x4320_rgb", (4320, 7680, 3), 3, lambda x, y: device_mod.sobel_edge_detection(x, 1, out=y)),
            ("c3_shape_gaussian_s2_r3", (4320, 7680, 3), 3, lambda x, y: device_mod.gaussian_blur(x, 2.0, 3, 1, out=y)),
            ("c3_shape_box_r3", (4320, 7680, 3), 3, lambda x, y: device_mod.box_blur(x, 3, 2, out=y)),
            ("c3_shape_gaussian_s5_r15", (4320, 7680, 3), 3, lambda x, y: device_mod.gaussian_blur(x, 5.0, 15, 1, out=y)),
            ("c3_shape_gaussian_s10_r31", (4320, 7680, 3), 3, lambda x, y: device_mod.gaussian_blur(x, 10.0, 31, 2, out=y))):
        try:
            xs = [torch.randint(0, 256, (h, w, c), dtype=torch.uint8, device="cuda", generator=g) for _ in range(nb)]
            ys = [torch.empty_like(t) for t in xs]
            out[name] = timed(lambda i: call(xs[i % nb], ys[i % nb]), h * w * c, h * w)
            del xs, ys
        except Exception as e:
            out[name] = {"error": repr(e)}
    try:    # c2: the radius sweep, one launch per radius on 4 rotating 64 MiB image pairs (512 MiB > L2)
        h, w, c, radii = C2["h"], C2["w"], C2["c"], C2["radii"]
        xs = [torch.randint(0, 256, (h, w, c), dtype=torch.uint8, device="cuda", generator=g) for _ in range(4)]
        ys = [torch.empty_like(t) for t in xs]

        def sweep(k):
            for i, r in enumerate(radii):
                b = (k * len(radii) + i) % 4
                device_mod.box_blur(xs[b], r, 2, out=ys[b])
        res = timed(sweep, len(radii) * h * w * c, len(radii) * h * w, reps=10)
        res["us_per_launch"] = res["us"] / len(radii)
        res["kernel"] = "gip_box_fused<4,true,*,8>"
        per_r = {}
        for r in (1, 3, 7, 8, 15, 16, 31):
            t = timed(lambda i: device_mod.box_blur(xs[i % 4], r, 2, out=ys[i % 4]), h * w * c, h * w, reps=8)
            per_r[f"r{r}"] = {"us": t["us"], "frac_of_hbm_peak": t["frac_of_hbm_peak"]}
        res["per_radius"] = per_r
        out["c2_box_sweep_r1_31_4096sq_rgba"] = res
        del xs, ys
    except Exception as e:
        out["c2_box_sweep_r1_31_4096sq_rgba"] = {"error": repr(e)}
    return out


def numpy_contract(torch, L, _lib):
    """The reference's actual plugin call: gpu_filters.<filter>(numpy array) -> dict with a new array (pageable memory
    both ways), wall clock; beside it the same call on pinned buffers through the C ABI."""
    import numpy as np
    import gpu_filters
    out = {}
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 256, size=(FH, FW, FC), dtype=np.uint8) for _ in range(8)]
    calls = {"gaussian": lambda f: gpu_filters.gaussian_blur(f, sigma=G_SIGMA, radius=G_RADIUS, level=2),
             "box": lambda f: gpu_filters.box_blur(f, radius=B_RADIUS, level=2),
             "sobel": lambda f: gpu_filters.sobel_edge_detection(f, level=S_LEVEL)}
    n = 32
    for k in FILTERS:
        calls[k](frames[0])
    t0 = time.perf_counter()
    for k in FILTERS:
        for i in range(n):
            calls[k](frames[i % 8])
    dt = time.perf_counter() - t0
    out["c4_frames"] = {"value": 3 * n * FRAME_PIX / dt / 1e6, "unit": "Mpix/s", "ms_per_call": dt / (3 * n) * 1e3,
                        "api": "gpu_filters.gaussian_blur / box_blur / sobel_edge_detection(ndarray), one 1080p RGB frame per call",
                        "h2d_bytes_per_call": FRAME_BYTES, "d2h_bytes_per_call": FRAME_BYTES}
    big = rng.integers(0, 256, size=(C2["h"], C2["w"], C2["c"]), dtype=np.uint8)
    gpu_filters.box_blur(big, radius=3, level=2)
    reps = 5
    t0 = time.perf_counter()
    for i in range(reps):
        gpu_filters.box_blur(big, radius=3 + i, level=2)
    ms_np = (time.perf_counter() - t0) / reps * 1e3
    hx = torch.from_numpy(big).pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    m = _lib.Metrics()
    _lib.check(L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), C2["w"], C2["h"], C2["c"], 1, 3, 2, ctypes.byref(m)))
    t0 = time.perf_counter()
    for i in range(reps):
        _lib.check(L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), C2["w"], C2["h"], C2["c"], 1, 3 + i, 2, ctypes.byref(m)))
    ms_pin = (time.perf_counter() - t0) / reps * 1e3
    out["c2_64MiB_box"] = {"numpy_ms_per_image": ms_np, "pinned_ms_per_image": ms_pin, "numpy_over_pinned": ms_np / ms_pin,
                           "numpy_Mpix/s": C2["h"] * C2["w"] / ms_np / 1e3, "pinned_Mpix/s": C2["h"] * C2["w"] / ms_pin / 1e3}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames of the stream (smaller values are for development only)")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / other configs / reference kernels")
    ap.add_argument("--no-c5", action="store_true", help="skip the c5 block (development)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from gpu_image_processing_b200 import _lib, bands, device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    placement = {"bound": False, "why": "GIP_BENCH_NO_BIND"}
    if os.environ.get("GIP_BENCH_NO_BIND") != "1":
        from gpu_image_processing_b200 import affinity
        placement = affinity.bind_to_device_numa(local)     # pinned buffers below land on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.load()
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    frames_total = args.frames
    peak, peak_src = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    # ---- c4: this rank's frames, resident in HBM ------------------------------------------------------------------
    lo, hi = bands.shard_range(frames_total, rank, world)
    n = hi - lo
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.empty((n, FH, FW, FC), dtype=torch.uint8, device="cuda")
    for i in range(0, n, 256):
        j = min(n, i + 256)
        x[i:j] = torch.randint(0, 256, (j - i, FH, FW, FC), dtype=torch.uint8, device="cuda", generator=g)
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream()
    calls = {"gaussian": lambda: device.gaussian_blur(x, G_SIGMA, G_RADIUS, 2, out=y),
             "box": lambda: device.box_blur(x, B_RADIUS, 2, out=y),
             "sobel": lambda: device.sobel_edge_detection(x, S_LEVEL, out=y)}

    def step(marks=None):
        for k in FILTERS:
            calls[k]()
            if marks is not None:
                ev = torch.cuda.Event(True)
                ev.record(stream)
                marks.append(ev)

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.gip_launch_count()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    marks = []
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    for _ in range(steps):
        step(marks)
    e1.record(stream)
    barrier()
    launches = L.gip_launch_count() - launches0
    per_ms = dict.fromkeys(FILTERS, 0.0)
    prev = e0
    for i, ev in enumerate(marks):
        per_ms[FILTERS[i % 3]] += prev.elapsed_time(ev)
        prev = ev
    red = max_over_ranks([e0.elapsed_time(e1)] + [per_ms[k] for k in FILTERS])
    ms_per_step = red[0] / steps
    per_ms = {k: red[1 + i] / steps for i, k in enumerate(FILTERS)}
    value = 3 * frames_total * FRAME_PIX / (ms_per_step / 1e3) / 1e6
    launches_all = int(max_over_ranks([float(launches)])[0]) * world if world > 1 else int(launches)

    # ---- end to end through the host-buffer entry points (what gpu_filters.* calls), pinned host memory ----------------
    e_lo, e_hi = bands.shard_range(min(E2E_FRAMES, frames_total), rank, world)
    ne = e_hi - e_lo
    e2e_steps = max(1, min(steps, 2))
    hx = torch.randint(0, 256, (ne, FH, FW, FC), dtype=torch.uint8).pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    m = _lib.Metrics()

    def e2e_step():
        _lib.check(L.gip_gaussian_blur_host(hx.data_ptr(), hy.data_ptr(), FW, FH, FC, ne, G_SIGMA, G_RADIUS, 3, ctypes.byref(m)))
        _lib.check(L.gip_box_blur_host(hx.data_ptr(), hy.data_ptr(), FW, FH, FC, ne, B_RADIUS, 2, ctypes.byref(m)))
        _lib.check(L.gip_sobel_host(hx.data_ptr(), hy.data_ptr(), FW, FH, FC, ne, S_LEVEL, ctypes.byref(m)))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks([time.perf_counter() - t0])[0]
    e2e_frames = min(E2E_FRAMES, frames_total)
    e2e_value = 3 * e2e_frames * FRAME_PIX * e2e_steps / e2e_s / 1e6
    del hx, hy
    L.gip_release_cache()
    del x, y
    torch.cuda.empty_cache()

    # ---- c5: one 32768 x 32768 RGB image, Gaussian r = 15: whole at N = 1, row bands + P2P halo rows at N > 1 ---------------
    c5 = None
    if not args.no_c5:
        H5, W5, C5c, r5, s5 = C5["h"], C5["w"], C5["c"], C5["radius"], C5["sigma"]
        reps = 5
        g5 = torch.Generator(device="cuda").manual_seed(100 + rank)
        bi = bands.BandedImage(H5, W5, C5c, r5, mode="p2p" if world > 1 else "copy")
        for i in range(0, bi.plan.rows, 2048):
            j = min(bi.plan.rows, i + 2048)
            bi.band[i:j] = torch.randint(0, 256, (j - i, W5, C5c), dtype=torch.uint8, device="cuda", generator=g5)
        bi.exchange()
        for _ in range(2):
            bi.filter("gaussian", sigma=s5, radius=r5, level=2)
        barrier()
        c0, c1 = torch.cuda.Event(True), torch.cuda.Event(True)
        c0.record(stream)
        for _ in range(reps):
            bi.filter("gaussian", sigma=s5, radius=r5, level=2)
        c1.record(stream)
        barrier()
        ms5 = max_over_ranks([c0.elapsed_time(c1) / reps])[0]
        gbs5 = 2.0 * H5 * W5 * C5c / ms5 / 1e6
        c5 = {"workload": "c5: one 32768x32768 RGB u8 image, Gaussian sigma=5 radius=15", "ms": ms5, "Mpix/s": H5 * W5 / ms5 / 1e3,
              "alg_GB/s": gbs5, "frac_of_hbm_peak_per_gpu": gbs5 / peak / world, "reps": reps, "n_gpus": world,
              "partition": "one launch set on one GPU" if world == 1 else f"{world} contiguous row bands of {bi.plan.rows} rows",
              "halo_mode": "none" if world == 1 else "p2p: the kernels read the neighbours' halo rows through CUDA-IPC peer pointers over NVLink (no exchange step, no staging copy)",
              "halo_rows_per_side": 0 if world == 1 else r5,
              "halo_bytes_per_rank": 0 if world == 1 else (bi.plan.rows_above + bi.plan.rows_below) * W5 * C5c,
              "halo_bytes_per_interior_rank": 0 if world <= 2 else 2 * r5 * W5 * C5c,
              "fp32": fp32_block(H5 * W5 * C5c / world, r5, ms5, None)}
        bi.finish()
        bi.close()
        torch.cuda.empty_cache()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        traffic = traffic_table()
        bytes_rank = (bands.shard_range(frames_total, 0, world)[1]) * FRAME_BYTES       # rank 0 holds the largest range
        per_filter, roof = {}, {}
        for k in FILTERS:
            ms = per_ms[k]
            gbs = 2.0 * frames_total * FRAME_BYTES / ms / 1e6                        # aggregate over ranks
            per_filter["c4_" + k] = {"ms": ms, "Mpix/s": frames_total * FRAME_PIX / ms / 1e3, "alg_GB/s": gbs,
                                     "frac_of_hbm_peak_per_gpu": gbs / world / peak, "share_of_step": ms / ms_per_step}
            t = traffic.get(k) or {}
            roof[k] = {"bound": "hbm", "kernel": KERNELS[k], "achieved": 2.0 * bytes_rank / ms / 1e6, "peak": peak, "unit": "GB/s",
                       "frac": 2.0 * bytes_rank / ms / 1e6 / peak, "frac_of_8TBs_nominal": 2.0 * bytes_rank / ms / 1e6 / 8000.0,
                       "algorithmic_bytes_per_launch": 2 * bytes_rank, "ms_per_launch": ms,
                       "traffic": (t.get("bytes_per_frame") * bytes_rank // FRAME_BYTES) if t.get("bytes_per_frame") else None,
                       "traffic_source": t.get("source"), "share_of_step": ms / ms_per_step}
        roof["gaussian"]["fp32"] = fp32_block(bytes_rank, G_RADIUS, per_ms["gaussian"], sm_mhz)
        if c5:
            c5["fp32"] = fp32_block(C5["h"] * C5["w"] * C5["c"] / world, C5["radius"], c5["ms"], sm_mhz)
        dominant = max(FILTERS, key=lambda k: per_ms[k])
        top = dict(roof[dominant])
        top["peak_source"] = peak_src
        top["step_frac"] = 3 * 2.0 * bytes_rank / ms_per_step / 1e6 / peak
        top["note"] = ("dominant kernel of the step by time; its own bound is the FP32 pipe (see fp32); step_frac = the step's total "
                       "algorithmic bytes (3 filters x 2 bytes per image byte) / step time / peak")
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": dict(CONFIG, frames=frames_total, frames_per_rank=hi - lo),
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": 3 * e2e_frames * FRAME_BYTES,
                    "d2h_bytes_per_step": 3 * e2e_frames * FRAME_BYTES, "steps": e2e_steps, "frames": e2e_frames,
                    "api": "gip_gaussian_blur_host / gip_box_blur_host / gip_sobel_host (C ABI behind gpu_filters.*), pinned host "
                           f"buffers, a {e2e_frames}-frame slice of the stream sharded over the ranks, per call H2D + kernels + D2H",
                    "cpu_placement_rank0": placement},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "roofline": top,
            "roofline_per_filter": roof,
            "filters": per_filter,
            "c5": c5,
        }
        if world == 1 and not args.no_extras:
            os.sched_setaffinity(0, all_cpus)                # the CPU baseline gets every core the box gives us
            line["cpu_baseline"], _ = cpu_port_run(1, 1)
            line["reference_cuda_same_gpu"] = reference_cuda_same_gpu(torch)
            try:
                line["filters"].update(other_configs(torch, device, peak))
            except Exception as e:
                line["filters"]["error"] = repr(e)
            try:
                line["e2e_numpy"] = numpy_contract(torch, L, _lib)
            except Exception as e:
                line["e2e_numpy"] = {"error": repr(e)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def fp32_block(nbytes, radius, ms, sm_mhz, sms=148):
    """The Gaussian's own roofline: exact parity with the reference costs (2r+1) dependent FMAs + 2 rounding adds per byte
    and pass, two passes: 4r+6 FP32 lane-operations per image byte, at 128 lanes per SM per clock."""
    ops = (4 * radius + 6) * float(nbytes)
    out = {"bound": "fp32", "ops_per_byte": 4 * radius + 6, "achieved": ops / ms / 1e9, "unit": "Tlane-op/s"}
    if sm_mhz:
        pk = 128.0 * sms * sm_mhz * 1e6 / 1e12
        out.update({"peak": pk, "frac": out["achieved"] / pk, "peak_source": f"128 FP32 lanes x {sms} SMs x {sm_mhz:.0f} MHz (sampled)"})
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Keep stdout for the one JSON line: libraries (NCCL's version banner, the reference's printf) write to fd 1,
    so point fd 1 at stderr for the run and keep a private duplicate of the real stdout for emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    quiet_stdout()
    sys.exit(main())
