"""``Session`` — drives one Engine for one simulation object (ours or the reference's).

Protocol per ``advance(n)`` call (SURVEY §8b "Step entry"):
  1. (re)compile sources/monitors into device ops           -> lowering.lower
  2. push coefficients if they changed, push the six host field arrays (honours user-set fields)
  3. tabulate amplitudes / phasors for the n steps on the host, run them on the device in chunks
  4. pull fields back into the SAME NumPy objects, fill the monitors' own result attributes
The host arrays stay the user-visible truth between calls, like the reference's in-place arrays.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from . import lowering
from .engine import Engine
from .grid import COMPONENTS, YeeGrid

_RECORD_POOL_BYTES = int(os.environ.get("PRISMO_B200_RECORD_POOL_MB", "512")) << 20
_MAX_TABLE_STEPS = 8192

_state = {"dtype": os.environ.get("PRISMO_B200_DTYPE", "float64"), "device": int(os.environ.get("PRISMO_B200_DEVICE", "0")),
          "flags": 0}


def configure(dtype: Optional[str] = None, device: Optional[int] = None, flags: Optional[int] = None,
              distributed: Optional[bool] = None) -> dict:
    """Engine-wide options: storage/arithmetic dtype ('float64' = parity default, 'float32' = fast), device,
    distributed (default True: under an initialised torch.distributed group with world_size > 1, 3-D simulations
    are slab-decomposed along x, one process per GPU)."""
    if distributed is not None:
        _state["distributed"] = bool(distributed)
    if dtype is not None:
        dt = np.dtype({"fp32": "float32", "fp64": "float64", "f32": "float32", "f64": "float64"}.get(dtype, dtype))
        if dt not in (np.float32, np.float64):
            raise ValueError(f"unsupported dtype {dtype}")
        _state["dtype"] = dt.name
    if device is not None:
        _state["device"] = int(device)
    if flags is not None:
        _state["flags"] = int(flags)
    return dict(_state)


def _device_count() -> int:
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 1


def _fingerprint(arrs):
    return tuple((id(a), a.shape, float(a.flat[0]), float(a.flat[-1]), float(a.sum())) for a in arrs)


class Session:
    def __init__(self, grid, dt: float, dtype=None, device: Optional[int] = None, flags: Optional[int] = None,
                 physics=None):
        """physics: None = parity mode (the reference's scheme, bug for bug).  A ``cpml.PMLParams`` (or True for
        defaults with thickness = the grid's boundary layers) selects the OPT-IN physics mode: stable Yee
        leap-frog + working CPML (3-D only; no reference numbers exist for it)."""
        self.grid = YeeGrid.like(grid)
        g = self.grid
        self.dt = float(dt)
        fl = _state["flags"] if flags is None else flags
        if physics:
            from . import _lib, cpml

            if g.is_2d:
                raise NotImplementedError("physics mode (stable Yee + CPML) is implemented for 3-D grids")
            fl |= _lib.FLAG_YEE
            params = physics if isinstance(physics, cpml.PMLParams) else cpml.PMLParams(thickness=g.pml_layers)
        from .distributed import SlabExecutor, active_world

        rank, world = active_world()
        self.distributed = bool(_state.get("distributed", True)) and world > 1 and g.is_3d and not physics
        if self.distributed:
            # one process per GPU under torchrun: this rank owns an x-slab (see distributed.py)
            dev = device if device is not None else rank % max(_device_count(), 1)
            self.engine = SlabExecutor(g, self.dt, dtype or _state["dtype"], dev, flags=fl, engine_cls=Engine)
        else:
            self.engine = Engine(3 if g.is_3d else 2, g.dimensions, g.spacing, self.dt,
                                 dtype=dtype or _state["dtype"], device=_state["device"] if device is None else device,
                                 flags=fl)
        if physics and params.thickness > 0:
            self.engine.set_cpml(params.thickness, cpml.coefficient_table(g.dimensions, g.spacing, self.dt, params))
        self._coef_sig = None

    def close(self):
        self.engine.close()

    # ---- coefficients ---------------------------------------------------------------------------------
    def set_coefficients(self, Ca, Cb, Da, Db) -> None:
        """Cell-centred update coefficients (core/solver.py:113-133); scalars or (nx,ny,nz) arrays."""
        if all(np.isscalar(a) for a in (Ca, Cb, Da, Db)):
            sig = ("u", float(Ca), float(Cb), float(Da), float(Db))
            if sig != self._coef_sig:
                self.engine.set_uniform_coeffs(Ca, Cb, Da, Db)
                self._coef_sig = sig
            return
        arrs = [np.asarray(a) for a in (Ca, Cb, Da, Db)]
        sig = _fingerprint(arrs)
        if sig == self._coef_sig:
            return
        if all(a.min() == a.max() for a in arrs):    # piecewise-constant everywhere -> the fused sweeps apply
            self.engine.set_uniform_coeffs(*[float(a.flat[0]) for a in arrs])
        else:
            self.engine.set_coeffs(*arrs)
        self._coef_sig = sig

    # ---- fields ------------------------------------------------------------------------------------------
    def push_fields(self, fields) -> None:
        for c in COMPONENTS:
            self.engine.upload(c, fields[c])

    def pull_fields(self, fields) -> None:
        for c in COMPONENTS:
            a = fields[c]
            if isinstance(a, np.ndarray):
                self.engine.download(c, a)
            else:                               # array-likes of foreign backends
                a[...] = self.engine.download(c)

    # ---- stepping ------------------------------------------------------------------------------------------
    @staticmethod
    def step_times(t0: float, dt: float, n: int) -> list:
        """Times seen by sources/monitors: accumulated ``t += dt`` (core/simulation.py:155-156)."""
        out, t = [], t0
        for _ in range(n):
            t += dt
            out.append(t)
        return out

    def advance(self, fields, sources, monitors, t0: float, dt: float, n: int, ades=()) -> float:
        """Run n full steps (H pass, E pass, sources, monitors).  Returns the new accumulated time."""
        if n <= 0:
            return t0
        eng = self.engine
        prog = lowering.lower(self.grid, sources, monitors, ades)
        eng.clear_ops()
        for op in prog.src_ops:
            eng.add_source_op(op)
        for op in prog.mon_ops:
            eng.add_monitor_op(op)
        for op in prog.ade_ops:
            eng.add_ade_op(op)
        for direction, lo, hi in prog.flux_ops:
            eng.add_flux_op(direction, lo, hi)
        chunk = n
        if prog.src_ops or prog.mon_ops or prog.flux_ops:
            chunk = min(chunk, _MAX_TABLE_STEPS)
            rec = prog.record_cells * 8
            if rec:
                chunk = max(1, min(chunk, _RECORD_POOL_BYTES // rec))
        self.push_fields(fields)
        done, t = 0, t0
        first = True
        while done < n:
            m = min(chunk, n - done)
            times = self.step_times(t, dt, m)
            if prog.src_ops or prog.mon_ops or prog.flux_ops:
                amp, ph = prog.tables(times, dt)
                eng.set_tables(m, amp, ph)
            if first:
                for b in prog.binders:
                    b.preload(eng)
                first = False
            eng.run(m)
            for b in prog.binders:
                b.collect(eng, times, dt, m)
            t = times[-1]
            done += m
        for b in prog.binders:
            b.finish(eng)
        self.pull_fields(fields)
        return t

    def half_step(self, fields, which: str) -> None:
        """MaxwellUpdater.update_magnetic_fields / update_electric_fields (core/solver.py:135-165)."""
        self.push_fields(fields)
        (self.engine.update_h if which == "H" else self.engine.update_e)()
        self.pull_fields(fields)
