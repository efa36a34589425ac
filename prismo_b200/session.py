"""``Session`` — drives one Engine for one simulation object (ours or the reference's).

Protocol per ``advance(n)`` call (SURVEY §8b "Step entry"):
  1. (re)compile sources/monitors into device ops           -> lowering.lower
  2. push coefficients if they changed; push the host field arrays THE USER MAY HAVE TOUCHED since the last call
  3. tabulate amplitudes / phasors for the n steps on the host (the next chunk's while the device runs the current
     one), run them on the device in chunks
  4. fill the monitors' own result attributes; the fields STAY ON THE DEVICE

Field residency.  The reference's ``ElectromagneticFields`` (core/fields.py:61-144) keeps its arrays in the dict
``_fields`` and every access — ``fields["Ez"]``, ``fields.Ez``, ``copy_fields_from``, ``zero_fields`` — goes through
that dict.  The session swaps the dict for a ``ResidentFields`` (a dict subclass) holding the same NumPy arrays:
  * after a run the host arrays are STALE; the first access to a component downloads it into the SAME array object,
  * an array that was handed out may have been written through, so it is uploaded again by the next advance(),
  * components nobody looked at never cross PCIe: a ``sim.step()`` loop or ``run(progress_callback=...)`` that does
    not read fields moves nothing, one that reads ``Ez`` moves Ez (down once per access point, up once per advance).
Containers without a ``_fields`` dict fall back to push-all / pull-all.
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


class ResidentFields(dict):
    """``fields._fields`` while a Session owns the device copy: component -> host mirror (the user's own array)."""

    def __init__(self, base, session):
        super().__init__(base)
        self.session = session
        self.stale = set()                      # host mirror older than the device copy
        self.touched = set(base)                # host mirror possibly newer than the device copy (handed out / replaced)

    def _raw(self, c):
        return dict.__getitem__(self, c)

    def __getitem__(self, c):
        a = dict.__getitem__(self, c)
        if c in self.stale:
            self.session._download_into(c, a)
            self.stale.discard(c)
        self.touched.add(c)
        return a

    def __setitem__(self, c, a):
        dict.__setitem__(self, c, a)
        self.stale.discard(c)
        self.touched.add(c)

    def get(self, c, default=None):
        return self[c] if c in self else default

    def values(self):
        return [self[c] for c in self.keys()]

    def items(self):
        return [(c, self[c]) for c in self.keys()]

    def sync_all(self):
        """Bring every host mirror up to date (the device copy is about to go away or be replaced)."""
        for c in list(self.stale):
            self.session._download_into(c, dict.__getitem__(self, c))
        self.stale.clear()

    def detach(self):
        self.sync_all()
        self.touched = set(self.keys())


def _fingerprint(arrs):
    """Identity + a strided sample (<= 64 Ki elements) of every coefficient array: O(1) per advance() instead of three
    full-grid reductions.  The reference builds Ca..Db once per MaxwellUpdater (core/solver.py:113-133) and never edits
    them; an in-place edit that misses the sample needs ``session.invalidate_coefficients()``."""
    out = []
    for a in arrs:
        flat = a.reshape(-1) if a.flags.c_contiguous else a.ravel()
        step = max(1, flat.size // 65536)
        sample = flat[::step]
        out.append((id(a), a.shape, float(flat[0]), float(flat[-1]), float(sample.sum()), float(sample.min()), float(sample.max())))
    return tuple(out)


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
        # (slabs need >= 4 planes each — the two-step sweep ships 4 —: smaller grids run whole on every rank)
        self.distributed = (bool(_state.get("distributed", True)) and world > 1 and g.is_3d and not physics
                            and g.dimensions[0] >= 4 * world)
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
        self._resident = None                   # the ResidentFields this session currently backs (if any)
        self.h2d_arrays = self.d2h_arrays = 0   # field arrays moved so far (tests / diagnostics)

    def close(self):
        if self._resident is not None:
            self._resident.detach()
            self._resident = None
        self.engine.close()

    def invalidate_coefficients(self) -> None:
        """Force the next advance() to re-send Ca, Cb, Da, Db (after editing them in place)."""
        self._coef_sig = None

    # ---- coefficients ---------------------------------------------------------------------------------
    def set_geometry(self, shapes, background=None, coords=None) -> None:
        """Paint a shape list into Ca, Cb, Da, Db ON THE DEVICE (geometry.py; geometry/shapes.py:71-99 +
        core/solver.py:113-133): nothing of full-grid size is built on the host or crosses PCIe.  background =
        (eps_r, mu_r, sigma_e, sigma_m) or a Material; coords = (x, y, z) cell coordinates (default: the grid's base
        coordinates, core/grid.py:194-197).  While a geometry is set, the host coefficient arrays of the updater are
        ignored (set_coefficients is a no-op) — clear_geometry() returns to them."""
        from . import geometry

        x, y, z = coords if coords is not None else geometry.cell_coordinates(self.grid)
        self.engine.rasterize(shapes, x, y, z, background)
        self._coef_sig = ("geometry",)

    def clear_geometry(self) -> None:
        self._coef_sig = None

    def set_coefficients(self, Ca, Cb, Da, Db) -> None:
        """Cell-centred update coefficients (core/solver.py:113-133); scalars or (nx,ny,nz) arrays."""
        if self._coef_sig == ("geometry",):
            return
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
        if all(a.min() == a.max() for a in arrs):    # constant everywhere -> the uniform fused sweeps apply
            self.engine.set_uniform_coeffs(*[float(a.flat[0]) for a in arrs])
        else:
            self.engine.set_coeffs(*arrs)
        self._coef_sig = sig

    # ---- fields ------------------------------------------------------------------------------------------
    def _download_into(self, c, a) -> None:
        if isinstance(a, np.ndarray):
            self.engine.download(c, a)
        else:                                   # array-likes of foreign backends
            a[...] = self.engine.download(c)
        self.d2h_arrays += 1

    def _bind(self, fields):
        """The ResidentFields behind `fields` (installing it on first use), or None for unknown containers."""
        d = getattr(fields, "_fields", None)
        if not isinstance(d, dict) or os.environ.get("PRISMO_B200_RESIDENT", "1") == "0":
            return None
        if isinstance(d, ResidentFields) and d.session is not self:
            d.detach()                          # another session's device copy: bring the host up to date first
            d.session = self
        elif not isinstance(d, ResidentFields):
            d = ResidentFields(d, self)
            fields._fields = d
        if self._resident is not d:
            if self._resident is not None:
                self._resident.detach()
            d.touched = set(d.keys())           # this engine has never seen these arrays
            self._resident = d
        return d

    def push_fields(self, fields) -> None:
        d = self._bind(fields)
        if d is None:
            for c in COMPONENTS:
                self.engine.upload(c, fields[c])
                self.h2d_arrays += 1
            return
        for c in COMPONENTS:
            if c in d.touched:
                self.engine.upload(c, d._raw(c))
                self.h2d_arrays += 1
        d.touched.clear()

    def pull_fields(self, fields) -> None:
        d = self._bind(fields)
        if d is None:
            for c in COMPONENTS:
                self._download_into(c, fields[c])
            return
        d.stale = set(COMPONENTS)               # downloaded lazily, on access

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
        tabled = bool(prog.src_ops or prog.mon_ops or prog.flux_ops)

        def plan(done, t):
            m = min(chunk, n - done)
            times = self.step_times(t, dt, m)
            return m, times, (prog.tables(times, dt) if tabled else None)

        done, t = 0, t0
        first = True
        nxt = plan(0, t0)
        while done < n:
            m, times, tabs = nxt
            if tabled:
                eng.set_tables(m, *tabs)
            if first:
                for b in prog.binders:
                    b.preload(eng)
                first = False
            eng.run(m)                          # asynchronous: the device works while the host tabulates the next chunk
            if done + m < n:
                nxt = plan(done + m, times[-1])
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
