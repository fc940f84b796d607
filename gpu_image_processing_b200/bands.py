"""Partitioning of the filter path across the GPUs of one box.

One process per GPU; `torch.distributed` is the plumbing (rendezvous, barriers, exchange of CUDA IPC
handles).  The data path has no collective:

  * image batches / frame streams (BASELINE config c4): every rank filters its own contiguous range
    of frames (`shard_range`) with one batched launch.  Nothing is exchanged.
  * one tall image (config c5): every rank owns a contiguous band of rows (`BandedImage`).  The stencil
    needs `halo` rows of the neighbouring bands (radius for the blurs, 1 for Sobel).  In mode "p2p" each
    rank maps its neighbours' band buffers with CUDA IPC and hands pointers INTO PEER MEMORY to the band
    kernels (gip_*_band in include/gip_b200.h): the kernels load the halo rows over NVLink while they
    stream their own rows, there is no separate exchange step and no staging copy.  Mode "copy" moves the
    halo rows with point-to-point send/recv instead (NCCL on GPU tensors; gloo on CPU tensors, which is
    what the CPU tests use to check the partition arithmetic).

The reference has no multi-GPU path (single device 0, default stream); this module is new work named by
BASELINE.json's north_star.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

KINDS = ("gaussian", "box", "sobel")


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-equal split of range(n): the first n % world ranks get one extra item."""
    if n < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def halo_rows(kind: str, radius: int = 1) -> int:
    if kind not in KINDS:
        raise ValueError(f"unknown filter {kind!r}")
    return 1 if kind == "sobel" else int(radius)


@dataclass
class BandPlan:
    """Rows [y0, y1) of an image of `height` rows and the halo rows that exist on either side."""
    height: int
    y0: int
    y1: int
    rows_above: int
    rows_below: int

    @property
    def rows(self) -> int:
        return self.y1 - self.y0


def plan_band(height: int, rank: int, world: int, halo: int) -> BandPlan:
    y0, y1 = shard_range(height, rank, world)
    return BandPlan(height, y0, y1, min(halo, y0), min(halo, height - y1))


class _DeviceBuffer:
    """cudaMalloc'ed memory (not the torch caching allocator: CUDA IPC exports whole allocations)."""

    def __init__(self, nbytes: int):
        p = ctypes.c_void_p()
        _lib.check(_lib.load().gip_device_alloc(nbytes, ctypes.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def tensor(self, shape) -> torch.Tensor:
        return torch.as_tensor(self, device="cuda").view(shape)

    def free(self):
        if self.ptr:
            _lib.load().gip_device_free(self.ptr)
            self.ptr = 0


class BandedImage:
    """One rank's band of a (height, width, channels) u8 image, plus the plumbing to filter it in place
    of the whole image.  Every rank constructs it with the same arguments."""

    def __init__(self, height: int, width: int, channels: int, halo: int, group=None, mode: str = "auto",
                 device: Optional[torch.device] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.height, self.width, self.channels, self.halo = height, width, channels, halo
        self.pitch = width * channels
        self.plan = plan_band(height, self.rank, self.world, halo)
        if self.plan.rows < halo and self.world > 1:
            raise ValueError("bands are shorter than the halo: use fewer ranks")
        self.cuda = torch.cuda.is_available() if device is None else torch.device(device).type == "cuda"
        if mode == "auto":
            mode = "p2p" if self.cuda and self.world > 1 else "copy"
        if mode == "p2p" and not self.cuda:
            raise ValueError("p2p needs CUDA devices")
        self.mode = mode
        self._bufs: List[_DeviceBuffer] = []
        self._peers = {}
        shape = (self.plan.rows, width, channels)
        if self.cuda and self.mode == "p2p":
            self._in = _DeviceBuffer(max(1, self.plan.rows * self.pitch)); self._bufs.append(self._in)
            self.band = self._in.tensor(shape)
        else:
            self.band = torch.empty(shape, dtype=torch.uint8, device="cuda" if self.cuda else "cpu")
        self.out = torch.empty_like(self.band)
        dev = self.band.device
        self.above = torch.empty((self.plan.rows_above, width, channels), dtype=torch.uint8, device=dev)
        self.below = torch.empty((self.plan.rows_below, width, channels), dtype=torch.uint8, device=dev)
        self._ptr_above = self._ptr_below = None
        if self.mode == "p2p":
            self._map_peers()

    # -- p2p: map the neighbours' band buffers -----------------------------------------------------
    def _map_peers(self):
        L = _lib.load()
        handle = (ctypes.c_uint8 * 64)()
        _lib.check(L.gip_ipc_export(self._in.ptr, handle))
        mine = (bytes(handle), self.plan.y0, self.plan.y1)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        for nb in (self.rank - 1, self.rank + 1):
            if 0 <= nb < self.world:
                h, y0, y1 = everyone[nb]
                p = ctypes.c_void_p()
                buf = (ctypes.c_uint8 * 64).from_buffer_copy(h)
                _lib.check(L.gip_ipc_open(buf, ctypes.byref(p)))
                self._peers[nb] = (p.value, y0, y1)
        if self.rank - 1 in self._peers and self.plan.rows_above:
            base, y0, y1 = self._peers[self.rank - 1]
            self._ptr_above = base + (y1 - y0 - self.plan.rows_above) * self.pitch   # its last rows_above rows
        if self.rank + 1 in self._peers and self.plan.rows_below:
            self._ptr_below = self._peers[self.rank + 1][0]                           # its first rows
        dist.barrier(group=self.group)

    # -- halo movement ------------------------------------------------------------------------------
    def exchange(self):
        """Make the neighbours' halo rows visible.  p2p: a barrier (the inputs must be complete before
        anyone reads them through the mapped pointers).  copy: send/recv of the halo rows."""
        if self.world == 1:
            return
        if self.mode == "p2p":
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            return
        ops = []
        up, down = self.rank - 1, self.rank + 1
        if up >= 0:
            ops.append(dist.P2POp(dist.isend, self.band[: self.halo].contiguous(), self._global(up), self.group))
            if self.plan.rows_above:
                ops.append(dist.P2POp(dist.irecv, self.above, self._global(up), self.group))
        if down < self.world:
            ops.append(dist.P2POp(dist.isend, self.band[self.plan.rows - self.halo:].contiguous(), self._global(down), self.group))
            if self.plan.rows_below:
                ops.append(dist.P2POp(dist.irecv, self.below, self._global(down), self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def _global(self, group_rank: int) -> int:
        return dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    # -- filtering ----------------------------------------------------------------------------------
    def filter(self, kind: str, sigma: float = 2.0, radius: int = 3, level: int = 1,
               compute: Optional[Callable] = None) -> torch.Tensor:
        """Rows [y0, y1) of filter(whole image).  Call exchange() first and finish() before `band` is rewritten (p2p
        mode: neighbours read this rank's band while their kernels run).  `compute(stitched, kind, params)` is a
        test hook for CPU tensors (the product path has no CPU implementation)."""
        if halo_rows(kind, radius) > self.halo:
            raise ValueError("halo too small for this radius")
        p = self.plan
        if not self.cuda:
            if compute is None:
                raise RuntimeError("CPU tensors: there is no CPU fallback (pass compute= in tests)")
            stitched = torch.cat([self.above, self.band, self.below], dim=0)
            full = compute(stitched, kind, dict(sigma=sigma, radius=radius, level=level),
                           at_top=p.y0 == 0, at_bottom=p.y1 == p.height)
            self.out.copy_(full[p.rows_above: p.rows_above + p.rows])
            return self.out
        L = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        if self.mode == "p2p":
            above, below = self._ptr_above, self._ptr_below
        else:
            above = self.above.data_ptr() if p.rows_above else None
            below = self.below.data_ptr() if p.rows_below else None
        args = (self.band.data_ptr(), above, below, self.out.data_ptr(), self.width, self.height, self.channels,
                p.y0, p.rows, p.rows_above, p.rows_below)
        if kind == "gaussian":
            rc = L.gip_gaussian_blur_band(*args, float(sigma), int(radius), 1 if level == 1 else 3, stream)
        elif kind == "box":
            rc = L.gip_box_blur_band(*args, int(radius), int(level), stream)
        elif kind == "sobel":
            rc = L.gip_sobel_band(*args, int(level), stream)
        else:
            raise ValueError(kind)
        _lib.check(rc)
        return self.out

    def finish(self):
        """Fence after filter() (p2p mode): filter() is stream-ordered and the neighbours' kernels read halo rows straight
        out of this rank's `band` over NVLink, so `band` must not be rewritten (next frame) until every rank's filter()
        has completed.  In a frame loop: write band -> exchange() -> filter() -> finish() -> write the next frame."""
        if self.cuda:
            torch.cuda.synchronize()
        if self.mode == "p2p" and self.world > 1:
            dist.barrier(group=self.group)

    def close(self):
        if self.mode == "p2p" and self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        L = _lib.load() if self._peers else None
        for ptr, _, _ in self._peers.values():
            L.gip_ipc_close(ptr)
        self._peers = {}
        self.band = self.out = None
        for b in self._bufs:
            b.free()
        self._bufs = []
