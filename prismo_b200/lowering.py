"""Lower source / monitor objects to device ops ("compile" a Simulation for the engine).

Works on duck-typed objects: either ``prismo_b200.sources/monitors`` descriptions or the reference's own
``prismo.sources/monitors`` instances (same public attributes).  Dispatch is by class name along the MRO, so
user subclasses of a stock class work; anything else raises ``NotImplementedError`` — there is no CPU path.

What a lowering produces (see ``Program``):
  * source ops   ``F[box] += amp(step) [* profile / divisor]``   + one host callable per amplitude column
  * monitor ops  sampled boxes (record and/or running DFT)      + one host callable per phasor column
  * binders      objects that copy device results back into the monitor's own result attributes

Reference semantics being reproduced (file:line under /root/reference/src/prismo):
  index boxes   core/grid.py:328-354, :383-513; sources/base.py:97-139; monitors/base.py:93-134
  PointSource   sources/point.py:47-73        PlaneWaveSource  sources/plane_wave.py:184-221
  TFSFSource    sources/tfsf.py:258-411       GaussianBeam     sources/gaussian.py:167-298
  ModeSource    sources/mode.py:84-361        FieldMonitor     monitors/field.py:111-143
  DFTMonitor    monitors/dft.py:108-160       FluxMonitor      monitors/flux.py:107-215
  ModeExpansion monitors/mode_monitor.py:135-212
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from .engine import AdeOp, MonitorOp, SourceOp
from .grid import COMPONENTS, YeeGrid
from . import postprocess

_EH = {  # (direction, polarization) -> (E component, H component)
    ("x", "y"): ("Ey", "Hz"), ("x", "z"): ("Ez", "Hy"), ("y", "x"): ("Ex", "Hz"),
    ("y", "z"): ("Ez", "Hx"), ("z", "x"): ("Ex", "Hy"), ("z", "y"): ("Ey", "Hx"),
}
ETA0 = np.sqrt((4 * np.pi * 1e-7) / 8.854187817e-12)     # sources/tfsf.py:286-290


@dataclass
class Program:
    grid: YeeGrid
    src_ops: list = field(default_factory=list)
    mon_ops: list = field(default_factory=list)
    amp_fns: list = field(default_factory=list)        # (t, dt) -> float
    phasor_fns: list = field(default_factory=list)     # t -> complex
    binders: list = field(default_factory=list)
    ade_ops: list = field(default_factory=list)
    flux_ops: list = field(default_factory=list)       # (direction, lo, hi): region-correct flux reductions
    _group: int = 0
    _phasor_cache: dict = field(default_factory=dict)

    # -- building blocks ---------------------------------------------------------------------------
    def amp(self, fn: Callable) -> int:
        self.amp_fns.append(fn)
        return len(self.amp_fns) - 1

    def phasors(self, fns, keys=None) -> int:
        """Columns of the phasor table for one monitor's frequency list.  Monitors with the SAME list (keys: one hashable
        per function, e.g. ("omega", w)) share their columns: the host evaluates each exp(-j w t) once per step instead of
        once per monitor and component (the per-step Python calls bound the API on small grids)."""
        if keys is not None:
            k = tuple(keys)
            hit = self._phasor_cache.get(k)
            if hit is not None:
                return hit
        first = len(self.phasor_fns)
        self.phasor_fns.extend(fns)
        if keys is not None:
            self._phasor_cache[k] = first
        return first

    def monitor_op(self, comp, box, record, n_freq, phasor_col) -> int:
        lo, hi = tuple(a for a, _ in box), tuple(b for _, b in box)
        self.mon_ops.append(MonitorOp(comp, lo, hi, bool(record), int(n_freq), int(phasor_col)))
        return len(self.mon_ops) - 1

    def source_ops(self, ops):
        """Add the ops of ONE source.  They share a group with earlier sources unless a box overlaps."""
        ops = [o for o in ops if all(h > l for l, h in zip(o.lo, o.hi))]
        cur = [o for o in self.src_ops if o.group == self._group]
        if any(_overlap(a, b) for a in ops for b in cur) or any(
                _overlap(a, b) for i, a in enumerate(ops) for b in ops[i + 1:]):
            self._group += 1
        for n, o in enumerate(ops):
            # ops of one source that overlap each other (never the case for stock sources) are serialised
            o.group = self._group
            if any(_overlap(o, p) for p in ops[:n]):
                self._group += 1
                o.group = self._group
        self.src_ops.extend(ops)

    # -- tables ----------------------------------------------------------------------------------------
    def tables(self, times, dt):
        """Evaluate every amplitude / phasor column at the given step times, one scalar call per step
        exactly as the reference's per-step calls do (keeps fp64 results bit-identical)."""
        n = len(times)
        amp = np.zeros((n, len(self.amp_fns)), dtype=np.float64)
        ph = np.zeros((n, len(self.phasor_fns)), dtype=np.complex128)
        for s, t in enumerate(times):
            for c, fn in enumerate(self.amp_fns):
                amp[s, c] = fn(t, dt)
            for c, fn in enumerate(self.phasor_fns):
                ph[s, c] = fn(t)
        return amp, ph

    @property
    def record_cells(self) -> int:
        return sum(int(np.prod([h - l for l, h in zip(o.lo, o.hi)])) for o in self.mon_ops if o.record)


def _overlap(a: SourceOp, b: SourceOp) -> bool:
    return a.component == b.component and all(al < bh and bl < ah for al, ah, bl, bh in zip(a.lo, a.hi, b.lo, b.hi))


def _kind(obj, table):
    for cls in type(obj).__mro__:
        if cls.__name__ in table:
            return cls.__name__
    return None


def _box(g: YeeGrid, comp, bounds):
    return g.component_box(comp, *bounds)


def _full_box(g: YeeGrid, comp):
    return tuple((0, n) for n in g.get_field_shape(comp))


def _src(comp, box, table, profile=None, divisor=1.0):
    return SourceOp(comp, tuple(a for a, _ in box), tuple(b for _, b in box), table, profile, divisor)


# ======================================================================================================
# sources
# ======================================================================================================
def _lower_point(p: Program, s):
    g = p.grid
    w = s.waveform
    p.source_ops([_src(s.component, _box(g, s.component, g.region_bounds(s.center, s.size)),
                       p.amp(lambda t, dt: w(t)))])


def _lower_plane_wave(p: Program, s):
    g = p.grid
    e, h = _EH[(s.direction.lower(), s.polarization.lower())]
    w, sign = s.waveform, s.direction_sign
    b = g.region_bounds(s.center, s.size)
    p.source_ops([
        _src(e, _box(g, e, b), p.amp(lambda t, dt: w(t) * sign)),
        _src(h, _box(g, h, b), p.amp(lambda t, dt: (w(t) / 377.0) * sign)),     # literal 377 (plane_wave.py:221)
    ])


def _lower_tfsf(p: Program, s):
    g = p.grid
    key = (s.direction, s.polarization)
    if key not in _EH:
        raise ValueError(f"Invalid direction-polarization combination: {key}")
    e, h = _EH[key]
    w, sign = s.waveform, s.direction_sign
    axis = "xyz".index(s.direction)
    if axis >= (3 if g.is_3d else 2):
        return                                          # z-propagation on a 2-D grid is a no-op (tfsf.py:398-411)
    ebox = _box(g, e, g.region_bounds(s.center, s.size))
    lo, hi = ebox[axis]
    plane = lo if sign > 0 else max(hi - 1, 0)          # entry surface (tfsf.py:228-256, :333-345)

    def slab(comp):
        box = list(_full_box(g, comp))                  # the WHOLE plane, not just the region (tfsf.py:338)
        box[axis] = (plane, plane + 1)
        return tuple(box)

    if sign > 0:
        fe = lambda t, dt: -w(t)                                        # E[i_min, :] -= e
        fh = lambda t, dt: -((w(t - 0.5 * dt) / ETA0) * sign)           # H[i_min, :] -= h*sign
    else:
        fe = lambda t, dt: w(t)                                         # E[i_max, :] += e
        fh = lambda t, dt: (w(t - 0.5 * dt) / ETA0) * sign              # H[i_max, :] += h*sign
    p.source_ops([_src(e, slab(e), p.amp(fe)), _src(h, slab(h), p.amp(fh))])


def _lower_gaussian(p: Program, s):
    g = p.grid
    e, h = _EH[(s.direction.lower(), s.polarization.lower())]
    x0, x1, y0, y1, z0, z1 = g.region_bounds(s.center, s.size)
    if g.is_2d and z0 == z1:
        z1 = z0 + 1
    d = s.direction.lower()
    c = s.center
    if d == "x":
        a, b = np.meshgrid(np.arange(y0, y1), np.arange(z0, z1), indexing="ij")
        r2 = (g.index_to_coord(1, a) - c[1]) ** 2
        if g.is_3d:
            r2 = r2 + (g.index_to_coord(2, b) - c[2]) ** 2
    elif d == "y":
        a, b = np.meshgrid(np.arange(x0, x1), np.arange(z0, z1), indexing="ij")
        r2 = (g.index_to_coord(0, a) - c[0]) ** 2
        if g.is_3d:
            r2 = r2 + (g.index_to_coord(2, b) - c[2]) ** 2
    else:
        a, b = np.meshgrid(np.arange(x0, x1), np.arange(y0, y1), indexing="ij")
        r2 = (g.index_to_coord(0, a) - c[0]) ** 2 + (g.index_to_coord(1, b) - c[1]) ** 2
    spatial = np.exp(-r2 / (s.beam_waist ** 2))         # x amplitude(t) on the device (gaussian.py:249)
    w = s.waveform
    tab = p.amp(lambda t, dt: w(t))
    bounds = g.region_bounds(s.center, s.size)
    ops = []
    for comp, div in ((e, 1.0), (h, 377.0)):
        box = _box(g, comp, bounds)
        shape = tuple(q - r for r, q in box)
        # "field[indices] += values.flat": NumPy broadcasts the flattened plane over the box
        # (gaussian.py:298); shapes that do not broadcast raise, exactly like the reference.
        prof = np.broadcast_to(np.asarray(spatial.flat), shape)
        ops.append(_src(comp, box, tab, np.ascontiguousarray(prof), div))
    p.source_ops(ops)


def _mode_profiles(g: YeeGrid, s):
    """Mode fields re-sampled on the source plane (sources/mode.py:84-196)."""
    have = getattr(s, "_mode_profile_Ex", None)
    if have is not None:                                 # a reference ModeSource already did it
        return {c: getattr(s, "_mode_profile_" + c) for c in COMPONENTS}
    from scipy.interpolate import RegularGridInterpolator

    m, c, sz = s.mode, s.center, s.size
    axes = {"z": (0, 1), "x": (1, 2), "y": (0, 2)}[s.axis]
    h = g.spacing
    lens = (len(m.x), len(m.y))
    lin = []
    for n, ax in enumerate(axes):
        if ax == 2 and g.is_2d:
            lin.append(np.linspace(0, 0, lens[n]))
        else:
            lin.append(np.linspace(c[ax] - sz[ax] / 2, c[ax] + sz[ax] / 2, max(int(sz[ax] / h[ax]), lens[n])))
    A, B = np.meshgrid(lin[0], lin[1], indexing="ij")
    pts = np.column_stack([A.ravel(), B.ravel()])

    def resample(f):
        re = RegularGridInterpolator((m.x, m.y), f.real, bounds_error=False, fill_value=0.0)
        im = RegularGridInterpolator((m.x, m.y), f.imag, bounds_error=False, fill_value=0.0)
        return re(pts).reshape(A.shape) + 1j * im(pts).reshape(A.shape)

    return {k: resample(getattr(m, k)) for k in COMPONENTS}


def _lower_mode(p: Program, s):
    g = p.grid
    if g.is_2d:
        raise IndexError("too many indices for array: ModeSource needs a 3-D grid (sources/mode.py:321-326)")
    prof = _mode_profiles(g, s)
    x0, x1, y0, y1, z0, z1 = g.region_bounds(s.center, s.size)
    if s.axis == "z":
        comps, tgt, fixed = ("Ex", "Ey", "Hz"), (x1 - x0, y1 - y0), 2
        want = [(x0, x1), (y0, y1), (z0, z0 + 1)]
    elif s.axis == "x":
        comps, tgt, fixed = ("Ey", "Ez", "Hx"), (y1 - y0, z1 - z0), 0
        want = [(x0, x0 + 1), (y0, y1), (z0, z1)]
    else:
        comps, tgt, fixed = ("Ex", "Ez", "Hy"), (x1 - x0, z1 - z0), 1
        want = [(x0, x1), (y0, y0 + 1), (z0, z1)]
    w, amp0, phase = s.waveform, s.amplitude, s.phase
    omega = 2 * np.pi * s.mode.frequency

    def amplitude(t, dt):                               # sources/mode.py:219-233
        return (amp0 * w.value(t) * np.exp(1j * (-omega * t + phase))).real

    tab = p.amp(amplitude)
    ops = []
    for n, comp in enumerate(comps):
        v = prof[comp].real
        if n == 2:
            v = v * s.sign
        if v.shape != tgt:
            from scipy.ndimage import zoom

            v = zoom(v, (tgt[0] / v.shape[0], tgt[1] / v.shape[1]), order=1)
        shape = g.get_field_shape(comp)
        if want[fixed][0] >= shape[fixed]:
            raise IndexError(f"index {want[fixed][0]} is out of bounds for axis {fixed} with size {shape[fixed]}")
        box = tuple((a, min(b, n_)) for (a, b), n_ in zip(want, shape))       # slices clip silently
        plane_shape = tuple(b - a for ax, (a, b) in enumerate(box) if ax != fixed)
        if plane_shape != v.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {plane_shape} {v.shape}")
        ops.append(_src(comp, box, tab, np.ascontiguousarray(np.expand_dims(v, fixed))))
    p.source_ops(ops)


_SOURCES = {"PointSource": _lower_point, "PlaneWaveSource": _lower_plane_wave, "TFSFSource": _lower_tfsf,
            "GaussianBeamSource": _lower_gaussian, "ModeSource": _lower_mode}


# ======================================================================================================
# monitors
# ======================================================================================================
class _Binder:
    """Copies device results into a monitor object's own attributes after each chunk."""

    def preload(self, engine):        # push existing DFT sums so that repeated runs keep accumulating
        pass

    def collect(self, engine, times, dt, n):
        pass

    def finish(self, engine):         # once, after the last chunk
        pass


def _patch_box(g: YeeGrid, comp, what):
    """The reference's placeholder region ``field[:10, :10]`` (monitors/dft.py:156-160)."""
    shape = g.get_field_shape(comp)
    if g.is_3d:
        # (10,10,nz) does not broadcast into the (n_f,10,10) accumulators: the reference raises here
        raise ValueError(f"{what}: non-broadcastable output operand with shape (10,10) doesn't match the "
                         f"broadcast shape (10,10,{shape[2]}) — the reference's placeholder monitor is 2-D only")
    if shape[0] < 10 or shape[1] < 10:
        raise ValueError(f"{what}: operands could not be broadcast together with shapes (10,10) "
                         f"({min(shape[0], 10)},{min(shape[1], 10)})")
    return ((0, 10), (0, 10))


class _FieldBinder(_Binder):
    def __init__(self, p: Program, m):
        g = p.grid
        self.m = m
        bounds = g.region_bounds(m.center, m.size)
        freqs = list(m.frequencies) if m.frequencies is not None else []
        self.freqs = freqs
        col = p.phasors([_phasor_scalar(f) for f in freqs], [("scalar", float(f)) for f in freqs]) if freqs else 0
        self.ids = {c: p.monitor_op(c, _box(g, c, bounds), m.time_domain, len(freqs), col) for c in m.components}

    def preload(self, engine):
        if self.freqs:
            for c, i in self.ids.items():
                engine.set_dft(i, np.stack([self.m._freq_data[c][f] for f in self.freqs]))

    def collect(self, engine, times, dt, n):
        m = self.m
        if m.time_domain:
            m._time_points.extend(times)
            for c, i in self.ids.items():
                m._time_data[c].extend(list(engine.records(i, n)))
        for c, i in self.ids.items():
            if self.freqs:
                acc = engine.dft(i)
                for k, f in enumerate(self.freqs):
                    m._freq_data[c][f][...] = acc[k]


def _phasor_scalar(freq):
    def fn(t):                                           # monitors/field.py:141-142
        omega = 2 * np.pi * freq
        return np.exp(-1j * omega * t)
    return fn


def _phasor_omega(omega):
    return lambda t: np.exp(-1j * omega * t)            # monitors/dft.py:129-132 (omega is an np.float64)


class _DFTBinder(_Binder):
    def __init__(self, p: Program, m):
        g = p.grid
        self.m = m
        n = len(m.omega)
        col = p.phasors([_phasor_omega(w) for w in m.omega], [("omega", float(w)) for w in m.omega])
        bounds = g.region_bounds(m.center, m.size)
        self.ids = {}
        for c in m.components:
            box = _box(g, c, bounds) if getattr(m, "region_correct", False) else _patch_box(g, c, "DFTMonitor")
            self.ids[c] = p.monitor_op(c, box, False, n, col)

    def preload(self, engine):
        for c, i in self.ids.items():
            engine.set_dft(i, self.m._dft_data[c])

    def collect(self, engine, times, dt, n):
        m = self.m
        if m._dt is None:
            m._dt = dt
        for c, i in self.ids.items():
            m._dft_data[c][...] = engine.dft(i)
        m._time_steps += n


class _PatchBinder(_Binder):
    """Six recorded corner patches per step (Flux / ModeExpansion placeholders)."""

    def __init__(self, p: Program, m, what, n_freq=0, col=0):
        g = p.grid
        self.m = m
        self.ids = {c: p.monitor_op(c, _patch_box(g, c, what), True, n_freq, col) for c in COMPONENTS}

    def patches(self, engine, n):
        return {c: engine.records(i, n) for c, i in self.ids.items()}


class _FluxRegionBinder(_Binder):
    """Extension (FluxMonitor(region_correct=True)): the monitor's REAL surface, in 2-D and 3-D.  Power through the
    box common to all six components is reduced on the device every step; the six DFTs run over the same box, so
    the reference's own ``get_frequency_domain_power`` (monitors/flux.py:228-291) works unchanged on the result."""

    def __init__(self, p: Program, m):
        g = p.grid
        self.m = m
        bounds = g.region_bounds(m.center, m.size)
        boxes = [_box(g, c, bounds) for c in COMPONENTS]
        box = tuple((max(b[a][0] for b in boxes), min(b[a][1] for b in boxes)) for a in range(len(boxes[0])))
        if any(hi <= lo for lo, hi in box):
            raise ValueError("FluxMonitor region is empty on the staggered grid")
        self.shape = tuple(hi - lo for lo, hi in box)
        self.nf = 0 if m.frequencies is None else len(m.omega)
        col = p.phasors([_phasor_omega(w) for w in m.omega], [("omega", float(w)) for w in m.omega]) if self.nf else 0
        self.ids = {c: p.monitor_op(c, box, False, self.nf, col) for c in COMPONENTS} if self.nf else {}
        p.flux_ops.append((m.direction, tuple(lo for lo, _ in box), tuple(hi for _, hi in box)))
        self.flux_id = len(p.flux_ops) - 1
        if self.nf:
            for c in COMPONENTS:
                name = "_dft_" + c.lower()
                if getattr(m, name, None) is None or getattr(m, name).shape != (self.nf,) + self.shape:
                    setattr(m, name, np.zeros((self.nf,) + self.shape, dtype=np.complex128))

    def preload(self, engine):
        for c, i in self.ids.items():
            engine.set_dft(i, getattr(self.m, "_dft_" + c.lower()))

    def collect(self, engine, times, dt, n):
        m = self.m
        dx, dy, dz = m._grid.spacing
        dA = {"x": dy * (dz if dz else 1.0), "y": dx * (dz if dz else 1.0), "z": dx * dy}[m.direction]
        m._power_flow_history.extend(float(v) * dA for v in engine.flux(self.flux_id, n))
        m._time_history.extend(times)

    def finish(self, engine):
        # the six DFT planes come down once per advance(), not once per chunk; until the next advance() they are also
        # still on the device, where postprocess.mode_coefficients reduces them without moving them
        for c, i in self.ids.items():
            getattr(self.m, "_dft_" + c.lower())[...] = engine.dft(i)
        if self.ids and hasattr(engine, "mode_overlap"):
            self.m._b200_device = (engine, [self.ids[c] for c in COMPONENTS], getattr(engine, "ops_epoch", 0))


class _FluxBinder(_PatchBinder):
    def __init__(self, p: Program, m):
        self.nf = 0 if m.frequencies is None else len(m.omega)
        col = p.phasors([_phasor_omega(w) for w in m.omega], [("omega", float(w)) for w in m.omega]) if self.nf else 0
        super().__init__(p, m, "FluxMonitor", self.nf, col)

    def preload(self, engine):
        if self.nf:
            for c, i in self.ids.items():
                engine.set_dft(i, getattr(self.m, "_dft_" + c.lower()))

    def collect(self, engine, times, dt, n):
        m = self.m
        rec = self.patches(engine, n)
        dx, dy, dz = m._grid.spacing
        dA = {"x": dy * dz, "y": dx * dz, "z": dx * dy}[m.direction]
        for s in range(n):
            m._power_flow_history.append(postprocess.patch_power(rec, s, m.direction, dA))
            m._time_history.append(times[s])
        if self.nf:
            for c, i in self.ids.items():
                getattr(m, "_dft_" + c.lower())[...] = engine.dft(i)


class _ModeExpansionBinder(_PatchBinder):
    def __init__(self, p: Program, m):
        super().__init__(p, m, "ModeExpansionMonitor")
        self.phasor = [_phasor_omega(w) for w in m.omega] if m.frequencies is not None else []

    def collect(self, engine, times, dt, n):
        m = self.m
        rec = self.patches(engine, n)
        for s in range(n):
            six = tuple(rec[c][s] for c in COMPONENTS)
            for i, mode in enumerate(m.modes):
                cf = postprocess.mode_overlap(six, mode, m.direction, 1.0, 1.0)   # dx = dy = 1 (mode_monitor.py:190)
                m._mode_coeffs_time[i].append(cf)
                for k, fn in enumerate(self.phasor):
                    m._mode_coeffs_freq[i][k] += cf * fn(times[s]) * dt
            m._time_points.append(times[s])


_MONITORS = {"FieldMonitor": _FieldBinder, "DFTMonitor": _DFTBinder, "FluxMonitor": _FluxBinder,
             "ModeExpansionMonitor": _ModeExpansionBinder}


class _AdeBinder(_Binder):
    """One dispersive medium (materials/ade.py:16-160) attached to one E component: one device recursion per
    Lorentz pole, or one for Drude / Debye.  State lives in the solver object between runs."""

    def __init__(self, p: Program, solver, component, mask):
        g = p.grid
        shape = g.get_field_shape(component)
        if tuple(solver.grid_shape) != tuple(shape):
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(solver.grid_shape)} {shape} "
                             f"(ADESolver.grid_shape must equal the {component} array shape)")
        self.solver = solver
        co = solver.coeffs
        lo, hi = (0,) * len(shape), tuple(shape)
        m = None if mask is None else np.broadcast_to(np.asarray(mask), shape)
        self.kind = 0 if "poles" in co else (1 if "omega_p" in co else 2)
        self.ids = []
        if self.kind == 0:
            for pole in co["poles"]:
                p.ade_ops.append(AdeOp(component, 0, lo, hi, pole["C0"], pole["C1"], pole["C2"], pole["C3"], m))
                self.ids.append(len(p.ade_ops) - 1)
        else:
            p.ade_ops.append(AdeOp(component, self.kind, lo, hi, co["C0"], co["C1"], 0.0, 0.0, m))
            self.ids.append(len(p.ade_ops) - 1)

    def preload(self, engine):
        s = self.solver
        if self.kind == 0:
            for n, i in enumerate(self.ids):
                engine.set_ade_state(i, 0, s.P_current[n])
                engine.set_ade_state(i, 1, s.P_previous[n])
        else:
            engine.set_ade_state(self.ids[0], 0, s.J_current if self.kind == 1 else s.P_current)

    def finish(self, engine):
        s = self.solver
        if self.kind == 0:
            for n, i in enumerate(self.ids):
                s.P_current[n] = engine.ade_state(i, 0)
                s.P_previous[n] = engine.ade_state(i, 1)
        elif self.kind == 1:
            s.J_current = engine.ade_state(self.ids[0], 0)
        else:
            s.P_current = engine.ade_state(self.ids[0], 0)


# ======================================================================================================
def lower(grid, sources, monitors, ades=()) -> Program:
    """Compile source and monitor objects for ``grid`` (any object exposing the reference's YeeGrid.spec)."""
    p = Program(YeeGrid.like(grid))
    for s in sources:
        if not getattr(s, "enabled", True):
            continue
        kind = _kind(s, _SOURCES)
        if kind is None:
            raise NotImplementedError(
                f"source type {type(s).__name__} cannot be lowered to the B200 engine (no CPU fallback); "
                f"supported: {sorted(_SOURCES)}")
        _SOURCES[kind](p, s)
    for m in monitors:
        kind = _kind(m, _MONITORS)
        if kind is None:
            raise NotImplementedError(
                f"monitor type {type(m).__name__} cannot be lowered to the B200 engine (no CPU fallback); "
                f"supported: {sorted(_MONITORS)}")
        if kind == "FluxMonitor" and getattr(m, "region_correct", False):
            p.binders.append(_FluxRegionBinder(p, m))
        else:
            p.binders.append(_MONITORS[kind](p, m))
    for solver, component, mask in ades:
        p.binders.append(_AdeBinder(p, solver, component, mask))
    return p
