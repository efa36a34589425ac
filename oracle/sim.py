"""Oracle restatement of the reference step loop, sources and monitors.  Test infrastructure only.

Follows /root/reference/src/prismo:
  core/simulation.py:147-164      step order: solver -> t += dt -> sources -> monitors   (a9)
  sources/point.py:47-73          PointSource                                             (a12)
  sources/plane_wave.py:184-221   PlaneWaveSource                                         (a13)
  sources/tfsf.py:258-411         TFSFSource (one whole plane, not a TF/SF box)           (a14)
  sources/gaussian.py:167-298     GaussianBeamSource (works in 2-D only)                  (a15)
  sources/mode.py:198-361         ModeSource (3-D only; needs waveform.value)             (a16)
  monitors/field.py:111-143       FieldMonitor                                            (a17)
  monitors/dft.py:108-160         DFTMonitor placeholder  ([:10,:10] corner patch)        (a18)
  monitors/flux.py:107-215        FluxMonitor placeholder                                 (a19)
  monitors/mode_monitor.py:135-212 + utils/mode_matching.py:41-171  ModeExpansionMonitor  (a20)
"""
from __future__ import annotations

import numpy as np

from . import kernels
from .grid import COMPONENTS, OGrid
from .waveforms import make_waveform

_PAIR = {  # (direction, polarization) -> (E comp, H comp)  plane_wave.py:162-172, tfsf.py:199-206
    ("x", "y"): ("Ey", "Hz"), ("x", "z"): ("Ez", "Hy"),
    ("y", "x"): ("Ex", "Hz"), ("y", "z"): ("Ez", "Hx"),
    ("z", "x"): ("Ex", "Hy"), ("z", "y"): ("Ey", "Hx"),
}


def _parse_dir(d):
    if d.startswith("+"):
        return d[1:], 1
    if d.startswith("-"):
        return d[1:], -1
    return d, 1


# ================================ sources ======================================================
class _Src:
    def init(self, grid: OGrid):
        self.grid = grid
        self.region = grid.region(self.center, self.size)
        self.box = {c: grid.box_slices(grid.component_box(c, self.region)) for c in COMPONENTS}


class PointSource(_Src):
    def __init__(self, position, component, waveform):
        self.center, self.size = position, (0, 0, 0)
        self.component, self.waveform = component, waveform

    def apply(self, F, t, dt):
        F[self.component][self.box[self.component]] += self.waveform(t)


class PlaneWaveSource(_Src):
    def __init__(self, center, size, direction, polarization, frequency, pulse=True,
                 pulse_width=None, amplitude=1.0, phase=0.0):
        self.center, self.size = center, size
        self.direction, self.sign = _parse_dir(direction)
        self.e, self.h = _PAIR[(self.direction.lower(), polarization.lower())]
        self.waveform = make_waveform(frequency, pulse, pulse_width, amplitude, phase)

    def apply(self, F, t, dt):
        a = self.waveform(t)
        F[self.e][self.box[self.e]] += a * self.sign
        F[self.h][self.box[self.h]] += (a / 377.0) * self.sign      # literal 377 (plane_wave.py:221)


class TFSFSource(_Src):
    def __init__(self, center, size, direction, polarization, frequency, pulse=True,
                 pulse_width=None, amplitude=1.0, phase=0.0):
        self.center, self.size = center, size
        self.direction, self.sign = _parse_dir(direction)
        self.e, self.h = _PAIR[(self.direction, polarization)]
        self.waveform = make_waveform(frequency, pulse, pulse_width, amplitude, phase)

    def init(self, grid):
        super().init(grid)
        b = grid.component_box(self.e, self.region)            # tfsf.py:228-256
        self.lo = [a for a, _ in b] + [0] * (3 - len(b))
        self.hi = [z - 1 for _, z in b] + [0] * (3 - len(b))

    def apply(self, F, t, dt):
        e = self.waveform(t)
        h = self.waveform(t - 0.5 * dt)
        eta0 = np.sqrt((4 * np.pi * 1e-7) / 8.854187817e-12)    # tfsf.py:286-290
        h = h / eta0
        ax = "xyz".index(self.direction)
        E, H = F[self.e], F[self.h]
        if ax >= E.ndim:                                        # z direction in 2-D: no-op (:398-411)
            return
        idx = [slice(None)] * E.ndim
        if self.sign > 0:                                       # whole plane, all other indices (:333-338)
            idx[ax] = self.lo[ax]
            E[tuple(idx)] -= e
            H[tuple(idx)] -= h * self.sign
        else:
            idx[ax] = self.hi[ax]
            E[tuple(idx)] += e
            H[tuple(idx)] += h * self.sign


class GaussianBeamSource(_Src):
    def __init__(self, center, size, direction, polarization, frequency, beam_waist, pulse=True,
                 pulse_width=None, amplitude=1.0, phase=0.0):
        self.center, self.size = center, size
        self.direction = direction
        self.e, self.h = _PAIR[(direction.lower(), polarization.lower())]
        self.w0 = beam_waist
        self.waveform = make_waveform(frequency, pulse, pulse_width, amplitude, phase)

    def apply(self, F, t, dt):
        g = self.grid
        amp = self.waveform(t)
        x0, x1, y0, y1, z0, z1 = g.region(self.center, self.size)
        if not g.is_3d and z0 == z1:
            z1 = z0 + 1
        d = self.direction.lower()
        if d == "x":                                            # gaussian.py:184-208
            a, b = np.meshgrid(np.arange(y0, y1), np.arange(z0, z1), indexing="ij")
            pa = g.index_to_coord(1, a)
            if g.is_3d:
                r2 = (pa - self.center[1]) ** 2 + (g.index_to_coord(2, b) - self.center[2]) ** 2
            else:
                r2 = (pa - self.center[1]) ** 2
        elif d == "y":                                          # :210-231
            a, b = np.meshgrid(np.arange(x0, x1), np.arange(z0, z1), indexing="ij")
            pa = g.index_to_coord(0, a)
            if g.is_3d:
                r2 = (pa - self.center[0]) ** 2 + (g.index_to_coord(2, b) - self.center[2]) ** 2
            else:
                r2 = (pa - self.center[0]) ** 2
        else:                                                   # :233-246
            a, b = np.meshgrid(np.arange(x0, x1), np.arange(y0, y1), indexing="ij")
            r2 = (g.index_to_coord(0, a) - self.center[0]) ** 2 + (g.index_to_coord(1, b) - self.center[1]) ** 2
        prof = amp * np.exp(-r2 / (self.w0 ** 2))
        # "+= values.flat" broadcasts the flattened plane over the box (:298); 3-D raises
        F[self.e][self.box[self.e]] += np.asarray(prof.flat)
        F[self.h][self.box[self.h]] += np.asarray((prof / 377.0).flat)


class ModeSource(_Src):
    """mode: object with Ex..Hz (2-D complex), x, y, frequency; waveform needs .value (F9)."""

    def __init__(self, center, size, mode, direction, waveform, amplitude=1.0, phase=0.0):
        self.center, self.size = center, size
        self.mode, self.waveform = mode, waveform
        self.amplitude, self.phase = amplitude, phase
        self.sign = +1 if direction[0] == "+" else -1
        self.axis = direction[-1].lower()

    def init(self, grid):
        from scipy.interpolate import RegularGridInterpolator

        super().init(grid)
        g, c, s, m = grid, self.center, self.size, self.mode
        t_axes = {"z": (0, 1), "x": (1, 2), "y": (0, 2)}[self.axis]     # mode.py:97-144
        d = (g.dx, g.dy, g.dz)
        lens = (len(m.x), len(m.y))
        lin = []
        for n, ax in enumerate(t_axes):
            if ax == 2 and not g.is_3d:
                lo = hi = 0
                cnt = lens[n]
            else:
                lo, hi = c[ax] - s[ax] / 2, c[ax] + s[ax] / 2
                cnt = max(int(s[ax] / d[ax]), lens[n])
            lin.append(np.linspace(lo, hi, cnt))
        A, B = np.meshgrid(lin[0], lin[1], indexing="ij")
        pts = np.column_stack([A.ravel(), B.ravel()])

        def interp(f):                                                   # mode.py:147-183
            re = RegularGridInterpolator((m.x, m.y), f.real, bounds_error=False, fill_value=0.0)
            im = RegularGridInterpolator((m.x, m.y), f.imag, bounds_error=False, fill_value=0.0)
            return re(pts).reshape(A.shape) + 1j * im(pts).reshape(A.shape)

        self.prof = {k: interp(getattr(m, k)) for k in COMPONENTS}

    def apply(self, F, t, dt):
        from scipy.ndimage import zoom

        wv = self.waveform.value(t)
        omega = 2 * np.pi * self.mode.frequency
        amp = (self.amplitude * wv * np.exp(1j * (-omega * t + self.phase))).real   # mode.py:219-233
        x0, x1, y0, y1, z0, z1 = self.grid.region(self.center, self.size)
        if self.axis == "z":                                             # :255-289
            comps, tgt = ("Ex", "Ey", "Hz"), (x1 - x0, y1 - y0)
            sl = (slice(x0, x1), slice(y0, y1), z0)
        elif self.axis == "x":                                           # :291-325
            comps, tgt = ("Ey", "Ez", "Hx"), (y1 - y0, z1 - z0)
            sl = (x0, slice(y0, y1), slice(z0, z1))
        else:                                                            # :327-361
            comps, tgt = ("Ex", "Ez", "Hy"), (x1 - x0, z1 - z0)
            sl = (slice(x0, x1), y0, slice(z0, z1))
        contrib = []
        for n, cname in enumerate(comps):
            v = amp * self.prof[cname].real
            if n == 2:
                v = v * self.sign
            contrib.append(v)
        if contrib[0].shape != tgt:
            zf = (tgt[0] / contrib[0].shape[0], tgt[1] / contrib[0].shape[1])
            contrib = [zoom(v, zf, order=1) for v in contrib]
        for cname, v in zip(comps, contrib):
            F[cname][sl] += v


# ================================ monitors =====================================================
class _Mon:
    def init(self, grid: OGrid):
        self.grid = grid
        self.region = grid.region(self.center, self.size)
        self.box = {c: grid.box_slices(grid.component_box(c, self.region)) for c in COMPONENTS}


class FieldMonitor(_Mon):
    def __init__(self, center, size, components="all", time_domain=True, frequencies=None):
        self.center, self.size = center, size
        self.components = {"all": list(COMPONENTS), "E": ["Ex", "Ey", "Ez"],
                           "H": ["Hx", "Hy", "Hz"]}.get(components, components) \
            if isinstance(components, str) else list(components)
        self.time_domain = time_domain
        self.frequencies = list(frequencies) if frequencies is not None else []
        self.time_points, self.time_data, self.freq_data = [], {c: [] for c in self.components}, {}

    def init(self, grid):
        super().init(grid)
        for c in self.components:
            shp = grid.component_box(c, self.region)
            shp = tuple(b - a for a, b in shp)
            self.freq_data[c] = {f: np.zeros(shp, dtype=np.complex128) for f in self.frequencies}

    def update(self, F, t, dt):
        if self.time_domain:
            self.time_points.append(t)
        for c in self.components:
            d = F[c][self.box[c]].copy()
            if self.time_domain:
                self.time_data[c].append(d)
            for f in self.frequencies:                                   # field.py:139-143
                omega = 2 * np.pi * f
                self.freq_data[c][f] += d * np.exp(-1j * omega * t) * dt


def _patch(a):
    """The placeholder extraction every non-Field monitor uses (dft.py:156-160)."""
    return a[:10, :10] if a.ndim >= 2 else a[:10]


class DFTMonitor(_Mon):
    def __init__(self, center, size, frequencies, components=None):
        self.center, self.size = center, size
        self.frequencies = np.array(frequencies)
        self.omega = 2 * np.pi * self.frequencies
        self.components = ["Ex", "Ey", "Ez"] if components is None else components
        self.dft, self.time_steps = {}, 0

    def init(self, grid):
        super().init(grid)
        for c in self.components:                                        # dft.py:83-106
            self.dft[c] = np.zeros((len(self.frequencies), 10, 10), dtype=np.complex128)

    def update(self, F, t, dt):
        for c in self.components:
            d = _patch(F[c])
            for i, w in enumerate(self.omega):
                self.dft[c][i] += d * np.exp(-1j * w * t) * dt
        self.time_steps += 1


class FluxMonitor(_Mon):
    def __init__(self, center, size, direction, frequencies=None):
        self.center, self.size = center, size
        self.direction = direction.lower()
        self.frequencies = frequencies
        self.power, self.times = [], []
        if frequencies is not None:
            self.omega = 2 * np.pi * np.array(frequencies)

    def init(self, grid):
        super().init(grid)
        if self.frequencies is not None:
            n = len(self.frequencies)
            self.dft = {c: np.zeros((n, 10, 10), dtype=np.complex128) for c in COMPONENTS}

    def _dA(self):
        dx, dy, dz = self.grid.spacing
        return {"x": dy * dz, "y": dx * dz, "z": dx * dy}[self.direction]

    def update(self, F, t, dt):
        Ex, Ey, Ez, Hx, Hy, Hz = (_patch(F[c]) for c in COMPONENTS)
        Sx = Ey * Hz - Ez * Hy                                           # flux.py:146-149
        Sy = Ez * Hx - Ex * Hz
        Sz = Ex * Hy - Ey * Hx
        S = {"x": Sx, "y": Sy, "z": Sz}[self.direction]
        self.power.append(float(np.sum(S) * self._dA()))
        self.times.append(t)
        if self.frequencies is not None:
            for i, w in enumerate(self.omega):
                ph = np.exp(-1j * w * t)
                for c, d in zip(COMPONENTS, (Ex, Ey, Ez, Hx, Hy, Hz)):
                    self.dft[c][i] += d * ph * dt

    def frequency_power(self):
        """flux.py:228-291."""
        out = []
        for i in range(len(self.frequencies)):
            Ex, Ey, Ez, Hx, Hy, Hz = (self.dft[c][i] for c in COMPONENTS)
            Sx = 0.5 * np.real(Ey * np.conj(Hz) - Ez * np.conj(Hy))
            Sy = 0.5 * np.real(Ez * np.conj(Hx) - Ex * np.conj(Hz))
            Sz = 0.5 * np.real(Ex * np.conj(Hy) - Ey * np.conj(Hx))
            S = {"x": Sx, "y": Sy, "z": Sz}[self.direction]
            out.append(float(np.sum(S) * self._dA()))
        return np.array(out)


def _resize(f, shape):
    """utils/mode_matching.py:_resize_to_shape — scipy.ndimage.zoom(order=1), complex handled by parts."""
    from scipy.ndimage import zoom

    if f.shape == shape:
        return f
    zf = [s / fs for s, fs in zip(shape, f.shape)]
    if np.iscomplexobj(f):
        return zoom(f.real, zf, order=1) + 1j * zoom(f.imag, zf, order=1)
    return zoom(f, zf, order=1)


def mode_power(mode, direction, dx=1.0, dy=1.0):
    """utils/mode_matching.py:134-171: abs(0.5*Re(sum(E x H*)_n)*dx*dy)."""
    d = direction.lower()
    if d == "x":
        S = mode.Ey * np.conj(mode.Hz) - mode.Ez * np.conj(mode.Hy)
    elif d == "y":
        S = mode.Ez * np.conj(mode.Hx) - mode.Ex * np.conj(mode.Hz)
    else:
        S = mode.Ex * np.conj(mode.Hy) - mode.Ey * np.conj(mode.Hx)
    return float(abs(0.5 * np.real(np.sum(S)) * dx * dy))


def mode_overlap(sim6, mode, direction, dx=1.0, dy=1.0):
    """utils/mode_matching.py:41-131."""
    Ex, Ey, Ez, Hx, Hy, Hz = sim6
    m = {c: getattr(mode, c) for c in COMPONENTS}
    if m["Ex"].shape != Ex.shape:
        m = {c: _resize(v, Ex.shape) for c, v in m.items()}
    d = direction.lower()
    if d == "x":
        Ss = Ey * np.conj(m["Hz"]) - Ez * np.conj(m["Hy"])
        Sm = m["Ey"] * np.conj(Hz) - m["Ez"] * np.conj(Hy)
    elif d == "y":
        Ss = Ez * np.conj(m["Hx"]) - Ex * np.conj(m["Hz"])
        Sm = m["Ez"] * np.conj(Hx) - m["Ex"] * np.conj(Hz)
    else:
        Ss = Ex * np.conj(m["Hy"]) - Ey * np.conj(m["Hx"])
        Sm = m["Ex"] * np.conj(Hy) - m["Ey"] * np.conj(Hx)
    ov = 0.5 * np.sum(Ss + Sm) * dx * dy
    p = mode_power(mode, direction, dx, dy)
    return complex(ov / p) if abs(p) > 1e-20 else complex(0.0)


class ModeExpansionMonitor(_Mon):
    def __init__(self, center, size, modes, direction="x", frequencies=None):
        self.center, self.size = center, size
        self.modes, self.direction, self.frequencies = modes, direction.lower(), frequencies
        self.coeffs_time = {i: [] for i in range(len(modes))}
        self.times = []
        if frequencies is not None:
            self.omega = 2 * np.pi * np.array(frequencies)
            self.coeffs_freq = {i: np.zeros(len(frequencies), dtype=complex) for i in range(len(modes))}

    def update(self, F, t, dt):
        six = tuple(_patch(F[c]) for c in COMPONENTS)
        for i, m in enumerate(self.modes):
            cf = mode_overlap(six, m, self.direction, 1.0, 1.0)          # dx=dy=1 (mode_monitor.py:190-191)
            self.coeffs_time[i].append(cf)
            if self.frequencies is not None:
                for k, w in enumerate(self.omega):
                    self.coeffs_freq[i][k] += cf * np.exp(-1j * w * t) * dt
        self.times.append(t)


# ================================ the loop =====================================================
class OSimulation:
    """Mirror of prismo.Simulation for the oracle (simulation.py:46-164)."""

    def __init__(self, size, resolution, pml_layers=10, courant_factor=0.9, materials=None,
                 dtype=np.float64):
        self.grid = g = OGrid(size, resolution, pml_layers)
        self.dt = g.time_step(courant_factor)
        self.F = {c: np.zeros(g.shape(c), dtype=dtype) for c in COMPONENTS}
        self.set_materials(materials)
        self.sources, self.monitors, self.ades = [], [], []
        self.step_count, self.current_time = 0, 0.0

    def set_materials(self, materials=None):
        dims = self.grid.dims
        m = materials or {}
        one, zero = np.ones(dims), np.zeros(dims)
        self.coeffs = kernels.coefficients(
            np.asarray(m.get("eps_rel", one)), np.asarray(m.get("mu_rel", one)),
            np.asarray(m.get("sigma_e", zero)), np.asarray(m.get("sigma_m", zero)), self.dt)

    def add_source(self, s):
        s.init(self.grid)
        self.sources.append(s)

    def add_monitor(self, m):
        m.init(self.grid)
        self.monitors.append(m)

    def step(self):
        kernels.step(self.F, self.coeffs, self.grid.spacing, self.grid.is_2d)
        self.step_count += 1
        self.current_time += self.dt            # accumulated, not n*dt (simulation.py:155-156)
        for s in self.sources:
            s.apply(self.F, self.current_time, self.dt)
        for m in self.monitors:
            m.update(self.F, self.current_time, self.dt)
        for a in self.ades:                     # what a user does after sim.step(): solver.update_polarization(E)
            a.update(self.F)

    def run_steps(self, n):
        for _ in range(n):
            self.step()
