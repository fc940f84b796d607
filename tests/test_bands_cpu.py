"""world_size-2/3 gloo tests (CPU) of the multi-GPU partitioning plumbing: row bands + halo exchange
(gpu_image_processing_b200/bands.py, mode "copy") and batch sharding.  The compute step is injected
(the oracle, test-only); what is under test is the partition arithmetic, halo sizes, clamp semantics at
the true image edges and the send/recv pattern."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpu_image_processing_b200 import bands
from tests import synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _oracle_compute(stitched, kind, params, at_top, at_bottom):
    """Filter the stitched rows as if they were a whole image.  Valid for the band's own rows because the
    halo is >= the stencil radius; rows at a true image edge keep the edge semantics."""
    from oracle import oracle as O
    a = stitched.numpy()
    if kind == "gaussian":
        out = O.gaussian_blur(a, params["sigma"], params["radius"], nthreads=1)
    elif kind == "box":
        out = O.box_blur(a, params["radius"], nthreads=1)
    else:
        out = O.sobel(a, params["level"], nthreads=1)
    return torch.from_numpy(out)


def _worker(rank, world, port, h, w, c, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = synth.uniform(h, w, c, seed=99)
        ok = True
        for kind, radius in (("gaussian", 5), ("box", 7), ("sobel", 1)):
            halo = bands.halo_rows(kind, radius)
            bi = bands.BandedImage(h, w, c, halo, mode="copy", device="cpu")
            assert (bi.plan.y0, bi.plan.y1) == bands.shard_range(h, rank, world)
            bi.band.copy_(torch.from_numpy(img[bi.plan.y0:bi.plan.y1]))
            bi.exchange()
            # the halos are exactly the neighbouring rows of the full image
            assert np.array_equal(bi.above.numpy(), img[bi.plan.y0 - bi.plan.rows_above:bi.plan.y0])
            assert np.array_equal(bi.below.numpy(), img[bi.plan.y1:bi.plan.y1 + bi.plan.rows_below])
            out = bi.filter(kind, sigma=2.5, radius=radius, level=1, compute=_oracle_compute).numpy()
            whole = _oracle_compute(torch.from_numpy(img), kind, dict(sigma=2.5, radius=radius, level=1), True, True).numpy()
            # Sobel zeroes the borders of what it is given: interior bands must not see an artificial border
            ok &= bool(np.array_equal(out, whole[bi.plan.y0:bi.plan.y1]))
            with pytest.raises(RuntimeError):
                bi.filter(kind, radius=radius)          # CPU tensors without the test hook: no fallback
        # batch sharding covers every frame exactly once
        lo, hi = bands.shard_range(4096, rank, world)
        t = torch.zeros(4096, dtype=torch.int32); t[lo:hi] = 1
        dist.all_reduce(t)
        ok &= bool((t == 1).all())
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_bands_and_batch_sharding_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, 61, 47, 3, results), nprocs=world, join=True)
    assert all(results.get(r) is True for r in range(world)), dict(results)


def test_shard_range_and_plans():
    for n in (0, 1, 7, 4096, 32768):
        for world in (1, 2, 3, 8):
            cuts = [bands.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    p = bands.plan_band(32768, 0, 8, 15)
    assert (p.y0, p.y1, p.rows_above, p.rows_below) == (0, 4096, 0, 15)
    p = bands.plan_band(32768, 7, 8, 15)
    assert (p.y0, p.y1, p.rows_above, p.rows_below) == (28672, 32768, 15, 0)
    assert bands.halo_rows("sobel", 9) == 1 and bands.halo_rows("box", 9) == 9
    with pytest.raises(ValueError):
        bands.halo_rows("median")
