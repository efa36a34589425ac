"""x-slab domain decomposition of the fused FDTD step across the GPUs of one box.

One process per GPU (``torchrun``); ``torch.distributed`` is only the plumbing (rendezvous + NCCL
send/recv of halo planes); every field update runs in libfdtd_b200.so on this rank's engine.

Decomposition: rank r owns total-grid planes [x0_r, x0_r + nx_r).  Both curls of the
reference scheme are forward differences (core/solver.py:178-300), so data only ever flows from rank r+1
to rank r.  For the fused single sweep the step-n inputs a rank needs from its right neighbour are the
neighbour's planes 0 and 1 at time n:   Ex, Ey, Ez, Hy, Hz of plane 0   and   Ey, Ez of plane 1
(7 planes; H+ of the ghost plane is recomputed locally).  They land in planes nx_r and nx_r + 1 of the
local arrays, which every array owns anyway (ghost + guard, see fdtd_kernels.cuh).

Overlap: the ghosts are consumed only by the sweep's LAST x-segment, while the planes a rank must send
are final as soon as its own post-step (sources) is done.  So each step is
    post halo send/recv (comm stream)  ||  sweep planes [0, nx_r - tail)        (compute stream)
    wait for the halo                  ->  sweep planes [nx_r - tail, nx_r), flip sets, sources + monitors
Results are bitwise identical to the single-GPU run: same kernel, same per-cell operation order.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def slab_range(nx: int, rank: int, world: int):
    """Planes [x0, x0 + n) of rank: contiguous, sizes differ by at most one."""
    base, rem = divmod(nx, world)
    n = base + (1 if rank < rem else 0)
    return rank * base + min(rank, rem), n


def plane_costs(nx: int, plane_cells: int, source_ops=(), monitor_ops=(), op_path_planes: float = 15.0):
    """Cost of every x-plane in units of "one plane of the sweep", for load-balanced slabs.

    A plane of the two-step sweep moves ~24 B per cell and step.  A DFT monitor read-modify-writes one complex128
    per cell, component and frequency each step (32 B, counted as 42.5: see below): a monitor op adds n_freq * 42.5 / 24
    * (its cells in the plane / plane_cells) plane-equivalents to every plane it covers; a recording op 8 B per cell.
    Every plane that carries any op also puts its x-segment on the op-carrying code path (~10 % slower over ~64 planes,
    plus the in-sweep injection): `op_path_planes`.  Ops are GLOBAL ops (x in global planes) with .lo / .hi boxes (SourceOp / MonitorOp)."""
    import numpy as np

    cost = np.ones(nx, dtype=np.float64)
    touched = np.zeros(nx, dtype=bool)
    for op in source_ops:
        touched[max(op.lo[0], 0):min(op.hi[0], nx)] = True
    for op in monitor_ops:
        a, b = max(op.lo[0], 0), min(op.hi[0], nx)
        if b <= a:
            continue
        touched[a:b] = True
        frac = (op.hi[1] - op.lo[1]) * (op.hi[2] - op.lo[2]) / float(plane_cells)
        # (calibrated on two 8-GPU per-rank breakdowns, profiles/r02_tuning.md §6: the rank with the 2 x 5-frequency DFT
        # plane behaves like 32.7 extra planes, the rank with the source plane like 15: the in-sweep running-DFT update is
        # a latency chain per cell, worth about 42 B per frequency rather than its 32 B of traffic)
        per_cell = 42.5 * getattr(op, "n_freq", 0) + (8.0 if getattr(op, "record", False) else 0.0)
        cost[a:b] += per_cell / 24.0 * frac
    cost[touched] += op_path_planes
    return cost


def balanced_slab_ranges(costs, world: int, min_planes: int = 4):
    """Contiguous x-slabs [(x0, n)] * world that minimise the LARGEST summed plane cost (the ranks run in lock
    step, so the most expensive slab sets the pace): bisection on the capacity + greedy fill.  Every slab keeps
    >= min_planes planes (the two-step sweep ships 4 of them).  Unit costs give slabs whose sizes differ by <= 1."""
    import numpy as np

    costs = np.asarray(costs, dtype=np.float64)
    nx = len(costs)
    if world < 1 or nx < world * min_planes:
        raise ValueError(f"{nx} planes cannot give {world} slabs of >= {min_planes} planes")
    cum = np.concatenate([[0.0], np.cumsum(costs)])

    def fill(cap):
        cuts = [0]
        for r in range(world):
            a = cuts[-1]
            hi_allowed = nx - (world - r - 1) * min_planes
            b = int(np.searchsorted(cum, cum[a] + cap * (1 + 1e-12), side="right")) - 1      # largest b: cost[a:b] <= cap
            b = min(max(b, a + min_planes), hi_allowed)
            cuts.append(b)
        return cuts

    lo, hi = cum[-1] / world, cum[-1]
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if fill(mid)[-1] >= nx:
            hi = mid
        else:
            lo = mid
    cuts = fill(hi)
    cuts[-1] = nx
    # the greedy fill leaves the slack in the last slab: hand planes back from the left while that lowers nothing
    # below the others (keeps unit-cost slabs within one plane of each other)
    for r in range(world - 1, 0, -1):
        while cuts[r] - cuts[r - 1] > min_planes and \
                (cum[cuts[r + 1]] - cum[cuts[r] - 1]) <= (cum[cuts[r]] - cum[cuts[r - 1]]):
            cuts[r] -= 1
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world)]


# (component, planes to ship): plane 0 of everything the ghost-plane H+ recompute reads, plane 1 of Ey/Ez
HALO_SPEC = (("Ex", 1), ("Ey", 2), ("Ez", 2), ("Hy", 1), ("Hz", 1))


class _CudaAlias:
    """Expose a raw device range to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerSlabRunner:
    """The production path: halo planes are DMA-pushed into the left neighbour's ghost planes through CUDA IPC
    mappings over NVLink, ordered by release/acquire flag words in peer memory; the whole step loop is enqueued
    by libfdtd_b200.so without host round trips (fdtd_slab_run).  torch.distributed only carries the one-off
    exchange of IPC handles."""

    def __init__(self, engine, rank: int, world: int, group=None):
        import torch.distributed as dist

        self.eng, self.rank, self.world = engine, rank, world
        blob = engine.ipc_export()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob, group=group)
        engine.ipc_connect(blobs[rank - 1] if rank > 0 else None, rank < world - 1)
        dist.barrier(group=group)              # nobody steps before everybody has reset its flags

    def run(self, n: int):
        self.eng.slab_run(n)

    def synchronize(self):
        self.eng.slab_sync()


class SlabStepper:
    """NCCL send/recv variant (and the gloo/CPU test vehicle).  Drives one engine (one x-slab) and its halo
    exchange from Python.  ``engine`` is a prismo_b200.Engine created with nx_global / x_offset, or any object
    with the same sweep / post_step / halo_tensors surface."""

    def __init__(self, engine, rank: int, world: int, tail_planes: int = 32, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.eng, self.rank, self.world, self.group = engine, rank, world, group
        self.nx = engine.dims[0]
        if world > 1 and self.nx < 3:
            raise ValueError("each slab needs at least 3 planes")
        self.tail = max(2, min(tail_planes, self.nx))     # >= 2: the head sweep reads up to plane head + 1
        self.cuda = hasattr(engine, "halo_ptrs")
        if self.cuda:
            self.compute = torch.cuda.Stream()
            self.comm = torch.cuda.Stream()
        self._alias_cache = {}

    # ---- halo buffers as torch tensors ------------------------------------------------------------------
    def _halo_tensors(self):
        """[(send tensor, recv tensor)] for the CURRENT buffer set."""
        if not self.cuda:
            return self.eng.halo_tensors(HALO_SPEC)
        torch = self.torch
        out = []
        esz = self.eng.dtype.itemsize
        typestr = "<f4" if esz == 4 else "<f8"
        for comp, planes in HALO_SPEC:
            first, ghost, plane_bytes = self.eng.halo_ptrs(comp)
            n = planes * plane_bytes // esz
            key = (first, ghost, n)
            if key not in self._alias_cache:
                self._alias_cache[key] = (torch.as_tensor(_CudaAlias(first, n, typestr), device="cuda"),
                                          torch.as_tensor(_CudaAlias(ghost, n, typestr), device="cuda"))
            out.append(self._alias_cache[key])
        return out

    def _post_exchange(self):
        """Send my planes 0/1 to rank-1, receive rank+1's into my ghost planes.  Returns the work handles."""
        dist = self.dist
        ops = []
        for send, recv in self._halo_tensors():
            if self.rank > 0:
                ops.append(dist.P2POp(dist.isend, send, self.rank - 1, self.group))
            if self.rank < self.world - 1:
                ops.append(dist.P2POp(dist.irecv, recv, self.rank + 1, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    # ---- stepping ---------------------------------------------------------------------------------------------
    def step(self):
        eng, torch = self.eng, self.torch
        if self.world == 1:
            eng.sweep(0, self.nx, True, self._s(self.compute) if self.cuda else 0)
            eng.post_step(self._s(self.compute) if self.cuda else 0)
            return
        if self.cuda:
            # the planes we send were finished by the previous post_step on the compute stream
            self.comm.wait_stream(self.compute)
            with torch.cuda.stream(self.comm):
                works = self._post_exchange()
            head = self.nx - self.tail
            eng.sweep(0, head, False, self._s(self.compute))
            with torch.cuda.stream(self.comm):
                for w in works:
                    w.wait()
            self.compute.wait_stream(self.comm)
            eng.sweep(head, self.nx, True, self._s(self.compute))
            eng.post_step(self._s(self.compute))
        else:
            works = self._post_exchange()
            head = self.nx - self.tail
            eng.sweep(0, head, False)
            for w in works:
                w.wait()
            eng.sweep(head, self.nx, True)
            eng.post_step()

    def run(self, n: int):
        for _ in range(n):
            self.step()

    def synchronize(self):
        if self.cuda:
            self.compute.synchronize()
            self.comm.synchronize()

    @staticmethod
    def _s(stream) -> int:
        return int(stream.cuda_stream)
