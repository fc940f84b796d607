"""Mint golden hashes from the reference's OWN kernels (oracle/_ref, built unmodified from
/root/reference/cuda_lib/src/image_filters.cu for sm_100a) on a B200.

    gpurun -- python -m tools.mint_golden            # writes gpurun_out/reference_hashes.json

The JSON is then committed as tests/golden/reference_hashes.json; tests/test_oracle.py checks on
the CPU that the oracle reproduces every hash, which pins the oracle to the reference itself.
Inputs are regenerated from (kind, shape, seed) by tests/synth.py, so only hashes are stored.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402


def _cases():
    out = []
    shapes = [(1, 1), (1, 7), (7, 1), (3, 3), (17, 33), (61, 40), (270, 481)]
    for c in (1, 3, 4):
        for (h, w) in shapes:
            for kind in ("uniform", "smooth"):
                if kind == "smooth" and h * w < 100:
                    continue
                base = dict(h=h, w=w, c=c, kind=kind, seed=1000 + 7 * h + w + c)
                for r, s in ((1, 0.8), (3, 2.0), (7, 3.0), (15, 5.0)):
                    if r > 3 and h * w < 100:
                        continue
                    for lvl in (1, 2):
                        out.append(dict(base, filter="gaussian", radius=r, sigma=s, level=lvl))
                for r in (1, 3, 5, 16, 31):
                    if r > 3 and h * w < 100:
                        continue
                    for lvl in (1, 2):
                        if lvl == 2 and r > 16:     # reference level-2 box is wrong above 16 (:489, :501)
                            continue
                        out.append(dict(base, filter="box", radius=r, sigma=0.0, level=lvl))
                for lvl in (1, 2):
                    out.append(dict(base, filter="sobel", radius=1, sigma=0.0, level=lvl))
    # the README benchmark shape (README.md:231-251), all three filters, level 1
    big = dict(h=2146, w=3239, c=3, kind="uniform", seed=4242)
    out.append(dict(big, filter="gaussian", radius=3, sigma=2.0, level=1))
    out.append(dict(big, filter="box", radius=5, sigma=0.0, level=1))
    out.append(dict(big, filter="sobel", radius=1, sigma=0.0, level=1))
    for d in out:
        d["name"] = "{filter}_L{level}_{kind}_{h}x{w}x{c}_r{radius}_s{sigma}".format(**d)
    return out


CASES = _cases()


def make_input(case):
    return synth.KINDS[case["kind"]](case["h"], case["w"], case["c"], seed=case["seed"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import torch
    from oracle import oracle as O

    assert torch.cuda.is_available(), "needs a GPU"
    assert O.ref_available(), "oracle/_ref/libref_image_filters.so missing"
    enum = {"gaussian": {1: 1, 2: 3}, "box": {1: 1, 2: 2}, "sobel": {1: 1, 2: 2}}
    gold, mismatches, nbytes = {}, [], 0
    for case in CASES:
        img = make_input(case)
        d_in = torch.from_numpy(img).cuda()
        d_out = torch.empty_like(d_in)
        torch.cuda.synchronize()
        rc, _ = O.ref_call(case["filter"], d_in.data_ptr(), d_out.data_ptr(), case["w"], case["h"], case["c"],
                           enum[case["filter"]][case["level"]], case["sigma"], case["radius"])
        torch.cuda.synchronize()
        assert rc == 0, (case["name"], rc)
        ref = d_out.cpu().numpy()
        if case["filter"] == "gaussian":
            ora = O.gaussian_blur(img, case["sigma"], case["radius"])
        elif case["filter"] == "box":
            ora = O.box_blur(img, case["radius"])
        else:
            ora = O.sobel(img, case["level"])
        gold[case["name"]] = {"sha256": sha(ref), "bytes": int(ref.size)}
        nbytes += ref.size
        if not np.array_equal(ref, ora):
            d = np.abs(ref.astype(int) - ora.astype(int))
            mismatches.append((case["name"], int(d.max()), float((d > 0).mean())))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    doc = {"source": "reference kernels (cuda_lib/src/image_filters.cu, unmodified) compiled nvcc -O3 sm_100a, run on "
                     + torch.cuda.get_device_name(0),
           "cases": gold}
    with open(os.path.join(ROOT, "gpurun_out", "reference_hashes.json"), "w") as f:
        json.dump(doc, f, indent=0, sort_keys=True)
    print(f"minted {len(gold)} cases, {nbytes} bytes; oracle mismatches: {len(mismatches)}")
    for m in mismatches[:40]:
        print("  MISMATCH", m)
    return 1 if mismatches else 0


if __name__ == "__main__":
    sys.exit(main())
