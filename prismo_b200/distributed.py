"""Slab-decomposed execution of a whole Simulation (sources, monitors included) under ``torchrun``.

Every rank builds the same Simulation; ``SlabExecutor`` gives ``Session`` the same surface as ``Engine`` but owns only
this rank's x-slab: global source / monitor ops are clipped to the slab (plus "ghost" copies of uniform source ops on
the three planes right of the slab, which the two-step sweep recomputes), the step loop is ``PeerSlabRunner`` (CUDA-IPC
halo push over NVLink) — or ``SlabStepper`` where the engine has no IPC (CPU test double under gloo) — and monitor
results are re-assembled from the ranks' pieces, so the reference's monitor objects end up with the same data on
every rank as in a single-GPU run.  3-D grids; uniform coefficients (one- and two-step fused sweeps) or heterogeneous
media (one-step sweep with one static ghost plane of Ca..Db from the right neighbour); dispersive-medium recursions
and region-correct flux sums are clipped to the slab like the monitors.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import Engine, MonitorOp, SourceOp
from .multigpu import PeerSlabRunner, SlabStepper, slab_range

_GATHER_FIELDS_MAX_BYTES = 1 << 30


def active_world():
    """(rank, world) of the default process group, or (0, 1) when torch.distributed is not in use."""
    try:
        import torch.distributed as dist
    except ImportError:
        return 0, 1
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


class SlabExecutor:
    def __init__(self, grid, dt, dtype, device, flags=0, engine_cls=None, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nx, ny, nz = grid.dimensions
        self.nxg = nx
        self.x0, self.nxl = slab_range(nx, self.rank, self.world)
        if self.nxl < 4:
            raise ValueError(f"{nx} planes over {self.world} ranks leaves slabs of < 4 planes")
        cls = engine_cls or Engine
        self.eng = cls(3, (self.nxl, ny, nz), grid.spacing, dt, dtype=dtype, device=device, nx_global=nx,
                       x_offset=self.x0, flags=flags)
        self.dtype, self.dt = self.eng.dtype, dt
        self._runner = None
        self._mon = []           # per global monitor op: (global op, local id or None, local x-range in the box)
        self._ade, self._flux = [], []
        self._tb2_ok = True
        self._agreed = False
        self._uploaded = False

    # ---- lifetime / passthrough ------------------------------------------------------------------------------
    def close(self):
        self.eng.close()

    @property
    def kernel_launches(self):
        return self.eng.kernel_launches

    def set_uniform_coeffs(self, *a):
        self.eng.set_uniform_coeffs(*a)
        self._het = False

    def set_coeffs(self, Ca, Cb, Da, Db):
        """Global cell-centred arrays (nx, ny, nz): this rank keeps its planes plus the right neighbour's first one (the
        E / H updates of the last local plane average coefficients across the cut, core/solver.py:458-533)."""
        hi = min(self.x0 + self.nxl + 1, self.nxg)
        self.eng.set_coeffs(*[np.ascontiguousarray(np.asarray(a)[self.x0:hi]) for a in (Ca, Cb, Da, Db)])
        self._het = True
        self._tb2_ok = False                      # heterogeneous media step one sweep at a time on every rank

    def rasterize(self, shapes, x, y, z=None, background=None):
        """Global cell coordinates: this rank paints its own planes plus the right neighbour's first one."""
        hi = min(self.x0 + self.nxl + 1, self.nxg)
        self.eng.rasterize(shapes, np.asarray(x, dtype=np.float64)[self.x0:hi], y, z, background)
        self._het = True
        self._tb2_ok = False

    def download_coeffs(self, which, planes=None):
        return self.eng.download_coeffs(which, planes)

    def _runner_(self):
        if self._runner is None:
            if hasattr(self.eng, "ipc_export"):
                self._runner = PeerSlabRunner(self.eng, self.rank, self.world, self.group)
            else:
                self._runner = SlabStepper(self.eng, self.rank, self.world, tail_planes=2, group=self.group)
        return self._runner

    # ---- fields ---------------------------------------------------------------------------------------------------
    def _planes(self, comp):
        return self.eng.field_shape(comp)[0]

    def upload(self, comp, array):
        a = np.asarray(array)
        self.eng.upload(comp, np.ascontiguousarray(a[self.x0:self.x0 + self._planes(comp)]))
        self._uploaded = True

    def download(self, comp, out=None):
        n = self._planes(comp)
        local = self.eng.download(comp)
        if out is None:
            shape = (self.nxg - (0 if comp in ("Ex", "Hy", "Hz") else 1),) + local.shape[1:]
            out = np.zeros(shape)
        out[self.x0:self.x0 + n] = local
        if out.nbytes <= _GATHER_FIELDS_MAX_BYTES:
            pieces = [None] * self.world
            self.dist.all_gather_object(pieces, (self.x0, local), group=self.group)
            for x0, piece in pieces:
                out[x0:x0 + piece.shape[0]] = piece
        return out

    # ---- ops ----------------------------------------------------------------------------------------------------------
    def clear_ops(self):
        self.eng.clear_ops()
        self._mon = []
        self._ade, self._flux = [], []
        self._tb2_ok = not getattr(self, "_het", False)
        self._agreed = False

    def _clip(self, comp, lo, hi):
        a, b = max(lo[0], self.x0), min(hi[0], self.x0 + self._planes(comp))
        return a, b

    def add_source_op(self, op: SourceOp):
        a, b = self._clip(op.component, op.lo, op.hi)
        if b > a:
            prof = None if op.profile is None else np.ascontiguousarray(op.profile[a - op.lo[0]:b - op.lo[0]])
            self.eng.add_source_op(SourceOp(op.component, (a - self.x0,) + tuple(op.lo[1:]), (b - self.x0,) + tuple(op.hi[1:]),
                                            op.table, prof, op.divisor, op.group))
        # ghost copy on the three planes right of the slab (two-step sweep recomputes the intermediate step there)
        if self.rank < self.world - 1:
            ga, gb = max(op.lo[0], self.x0 + self.nxl), min(op.hi[0], self.x0 + self.nxl + 3)
            if gb > ga:
                if op.profile is not None:
                    self._tb2_ok = False           # profiled ghost ops: fall back to the one-step sweep
                elif hasattr(self.eng, "ipc_export") or getattr(self.eng, "supports_ghost_ops", False):
                    self.eng.add_source_op(SourceOp(op.component, (ga - self.x0,) + tuple(op.lo[1:]),
                                                    (gb - self.x0,) + tuple(op.hi[1:]), op.table, None, 1.0, op.group,
                                                    ghost=True))

    def add_monitor_op(self, op: MonitorOp) -> int:
        a, b = self._clip(op.component, op.lo, op.hi)
        op.shape = tuple(h - l for l, h in zip(op.lo, op.hi))
        local = None
        if b > a:
            local = self.eng.add_monitor_op(MonitorOp(op.component, (a - self.x0,) + tuple(op.lo[1:]),
                                                      (b - self.x0,) + tuple(op.hi[1:]), op.record, op.n_freq, op.phasor_col))
        self._mon.append((op, local, (a - op.lo[0], b - op.lo[0])))
        return len(self._mon) - 1

    def add_ade_op(self, op) -> int:
        """Cell-local recursion: each rank keeps the part of the box that lies on its planes."""
        from .engine import AdeOp

        a, b = self._clip(op.component, op.lo, op.hi)
        op.shape = tuple(h - l for l, h in zip(op.lo, op.hi))
        local = None
        if b > a:
            mask = None if op.mask is None else np.ascontiguousarray(np.asarray(op.mask)[a - op.lo[0]:b - op.lo[0]])
            local = self.eng.add_ade_op(AdeOp(op.component, op.kind, (a - self.x0,) + tuple(op.lo[1:]),
                                              (b - self.x0,) + tuple(op.hi[1:]), op.c0, op.c1, op.c2, op.c3, mask))
        self._ade.append((op, local, (a - op.lo[0], b - op.lo[0])))
        self._tb2_ok = False
        return len(self._ade) - 1

    def ade_state(self, idx, which=0):
        op, local, (a, b) = self._ade[idx]
        mine = self.eng.ade_state(local, which) if local is not None else None
        pieces = [None] * self.world
        self.dist.all_gather_object(pieces, (a, b, mine), group=self.group)
        out = np.zeros(op.shape)
        for pa, pb, arr in pieces:
            if arr is not None:
                out[pa:pb] = arr
        return out

    def set_ade_state(self, idx, which, values):
        op, local, (a, b) = self._ade[idx]
        if local is not None:
            self.eng.set_ade_state(local, which, np.ascontiguousarray(np.asarray(values)[a:b]))

    def add_flux_op(self, direction, lo, hi) -> int:
        """Power through a box: per-rank partial sums over the planes it owns (all six components must exist there)."""
        a, b = max(lo[0], self.x0), min(hi[0], self.x0 + min(self._planes(c) for c in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")))
        local = self.eng.add_flux_op(direction, (a - self.x0,) + tuple(lo[1:]), (b - self.x0,) + tuple(hi[1:])) if b > a else None
        self._flux.append(local)
        self._tb2_ok = False
        return len(self._flux) - 1

    def flux(self, idx, steps):
        local = self._flux[idx]
        mine = self.eng.flux(local, steps) if local is not None else np.zeros(steps)
        parts = [None] * self.world
        self.dist.all_gather_object(parts, mine, group=self.group)
        out = np.zeros(steps)
        for p in parts:                            # fixed rank order: every rank gets the same bits
            out = out + p
        return out

    def set_tables(self, n_steps, amp=None, phasors=None):
        self.eng.set_tables(n_steps, amp, phasors)

    # ---- stepping -------------------------------------------------------------------------------------------------------
    def run(self, n):
        if not self._agreed:
            # every rank must pair steps identically (the halo protocol counts exchanges): the two-step sweep is
            # used only if NO rank has a reason to fall back to the one-step sweep
            votes = [None] * self.world
            self.dist.all_gather_object(votes, self._tb2_ok, group=self.group)
            if hasattr(self.eng, "set_option"):
                self.eng.set_option("tb2", 1 if all(votes) else 0)
            self._agreed = True
        r = self._runner_()
        if self._uploaded:
            # the right neighbour pushes its halo into our ghost planes as soon as ITS run starts: every rank's
            # uploads must be complete first (fdtd_upload_field in include/fdtd_b200.h)
            self.dist.barrier(group=self.group)
            self._uploaded = False
        r.run(n)
        r.synchronize()

    def update_h(self):
        raise NotImplementedError("half steps are not available on a slab-decomposed simulation")

    update_e = update_h

    # ---- monitor read-out: assemble the ranks' pieces along x -------------------------------------------------------------
    def _assemble(self, idx, lead, fetch):
        op, local, (a, b) = self._mon[idx]
        mine = fetch(local) if local is not None else None
        pieces = [None] * self.world
        self.dist.all_gather_object(pieces, (a, b, mine), group=self.group)
        out = np.zeros(lead + op.shape, dtype=np.complex128 if fetch == self.eng.dft else np.float64)
        ax = len(lead)
        for pa, pb, arr in pieces:
            if arr is not None:
                sl = [slice(None)] * out.ndim
                sl[ax] = slice(pa, pb)
                out[tuple(sl)] = arr
        return out

    def records(self, idx, steps):
        return self._assemble(idx, (steps,), lambda i: self.eng.records(i, steps))

    def dft(self, idx):
        op = self._mon[idx][0]
        return self._assemble(idx, (op.n_freq,), self.eng.dft)

    def set_dft(self, idx, values):
        op, local, (a, b) = self._mon[idx]
        if local is not None:
            self.eng.set_dft(local, np.ascontiguousarray(np.asarray(values)[:, a:b]))
