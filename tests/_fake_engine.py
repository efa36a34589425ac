"""NumPy test double for prismo_b200.engine.Engine — interprets the device op descriptors on the CPU.

TEST INFRASTRUCTURE.  It exists so that the host logic (lowering of sources/monitors to ops, tables,
chunking, write-back into monitor objects, the prismo plugin) can be verified without a GPU against the
real reference.  It implements exactly the op semantics documented in include/fdtd_b200.h, with the oracle
kernels standing in for the CUDA field update.  The product never imports this.
"""
from __future__ import annotations

import numpy as np

from oracle import kernels
from prismo_b200.grid import SHORT_AXES

COMPONENTS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


def shape_dict(sh):
    """A geometry object (mirror or reference class) as the rasterisation oracle's dict."""
    m = getattr(sh, "material", None)
    mat = {} if m is None else dict(eps_r=(tuple(m.epsilon_r) if np.ndim(m.epsilon_r) else m.epsilon_r), mu_r=m.mu_r,
                                    sigma_e=getattr(m, "sigma_e", 0.0), sigma_m=getattr(m, "sigma_m", 0.0))
    kind = type(sh).__name__
    if hasattr(sh, "operation"):
        d = dict(kind="group", operation=sh.operation, shapes=[shape_dict(q) for q in sh.shapes])
        if m is None:
            d.update({k: v for k, v in shape_dict(sh.shapes[0]).items() if k in ("eps_r", "mu_r", "sigma_e", "sigma_m")})
        return dict(d, **mat)
    if kind == "Box":
        return dict(kind="box", center=tuple(sh.center), size=tuple(sh.size), **mat)
    if kind == "Sphere":
        return dict(kind="sphere", center=tuple(sh.center), radius=sh.radius, **mat)
    if kind == "Cylinder":
        return dict(kind="cylinder", center=tuple(sh.center), radius=sh.radius, height=sh.height, axis=sh.axis, **mat)
    if kind == "Polygon":
        return dict(kind="polygon", vertices=np.asarray(sh.vertices), z_min=sh.z_min, z_max=sh.z_max, **mat)
    raise TypeError(kind)


class FakeEngine:
    instances = []

    def __init__(self, ndim, dims, spacing, dt, dtype="float64", device=0, nx_global=None, x_offset=0, flags=0):
        self.ndim, self.dims, self.spacing, self.dt = ndim, tuple(dims), tuple(spacing), dt
        self.dtype = np.dtype(dtype)
        self.F = {c: np.zeros(self.field_shape(c), dtype=np.float64) for c in COMPONENTS}
        full = self.dims if ndim == 3 else (self.dims[0], self.dims[1], 1)
        eps0, mu0 = 8.854187817e-12, 4 * np.pi * 1e-7
        self.coeffs = [np.full(full, v) for v in (1.0, dt / eps0, 1.0, dt / mu0)]
        self.src, self.mon, self.ade, self.fluxes = [], [], [], []
        self.cursor, self.n_tab = 0, 0
        self.kernel_launches = 0
        self.uploads = 0
        FakeEngine.instances.append(self)

    def close(self):
        pass

    def field_shape(self, comp):
        n = list(self.dims[: self.ndim])
        for ax in SHORT_AXES[comp]:
            if ax < self.ndim:
                n[ax] -= 1
        return tuple(n)

    def set_uniform_coeffs(self, ca, cb, da, db):
        full = self.dims if self.ndim == 3 else (self.dims[0], self.dims[1], 1)
        self.coeffs = [np.full(full, float(v)) for v in (ca, cb, da, db)]

    def set_coeffs(self, Ca, Cb, Da, Db):
        self.coeffs = [np.array(a, dtype=np.float64).reshape(self.dims if self.ndim == 3 else self.dims[:2] + (1,))
                       for a in (Ca, Cb, Da, Db)]

    def rasterize(self, shapes, x, y, z=None, background=None):
        """fdtd_rasterize: the rasterisation oracle stands in for k_rasterize (tests/test_raster.py pins both)."""
        from oracle import raster
        from prismo_b200 import geometry as G

        bg = G.background_values(background)
        bgt = ((bg[0] if bg[0] == bg[1] == bg[2] else tuple(bg[:3])), bg[3], bg[4], bg[5])
        self._painted = raster.coefficient_arrays([shape_dict(s) for s in shapes], x, y, z, self.dt, bgt)
        self._set_painted(len(x))

    def _set_painted(self, planes):
        assert planes == self.dims[0]
        self.coeffs = list(self._painted)

    def download_coeffs(self, which, planes=None):
        idx = {"Ca": 0, "Cb": 1, "Cbx": 1, "Da": 2, "Db": 3, "Cby": 4, "Cbz": 5}.get(which, which)
        Ca, Cb, Da, Db = self.coeffs
        cb3 = Cb if isinstance(Cb, (tuple, list)) else (Cb, None, None)
        a = (Ca, cb3[0], Da, Db, cb3[1], cb3[2])[idx]
        if a is None:
            raise RuntimeError("coefficient array is not set")
        return np.array(a[: (planes or self.dims[0])])

    def upload(self, comp, a):
        a = np.asarray(a)
        if a.shape != self.field_shape(comp):
            raise ValueError("shape")
        self.F[comp][...] = a
        self.uploads += 1

    def download(self, comp, out=None):
        if out is None:
            return self.F[comp].copy()
        out[...] = self.F[comp]
        return out

    def clear_ops(self):
        self.src, self.mon, self.ade, self.fluxes = [], [], [], []

    def add_source_op(self, op):
        self.src.append(op)

    def add_monitor_op(self, op):
        op.shape = tuple(h - l for l, h in zip(op.lo, op.hi))
        self.mon.append(dict(op=op, rec=[], dft=np.zeros((op.n_freq,) + op.shape, dtype=np.complex128)))
        return len(self.mon) - 1

    def add_flux_op(self, direction, lo, hi):
        self.fluxes.append(dict(d=direction, sl=tuple(slice(a, b) for a, b in zip(lo, hi)), out=[]))
        return len(self.fluxes) - 1

    def flux(self, i, steps):
        return np.array(self.fluxes[i]["out"][:steps])

    def _run_flux(self):
        for f in self.fluxes:
            ex, ey, ez, hx, hy, hz = (self.F[c][f["sl"]] for c in COMPONENTS)
            s = {"x": ey * hz - ez * hy, "y": ez * hx - ex * hz, "z": ex * hy - ey * hx}[f["d"]]
            f["out"].append(float(np.sum(s)))

    def add_ade_op(self, op):
        op.shape = tuple(h - l for l, h in zip(op.lo, op.hi))
        self.ade.append(dict(op=op, cur=np.zeros(op.shape), prev=np.zeros(op.shape)))
        return len(self.ade) - 1

    def ade_state(self, i, which=0):
        return self.ade[i]["cur" if which == 0 else "prev"].copy()

    def set_ade_state(self, i, which, v):
        self.ade[i]["cur" if which == 0 else "prev"][...] = v

    def _run_ade(self):
        for a in self.ade:
            o = a["op"]
            e = self.F[o.component][self._sl(o)]
            if o.mask is not None:
                e = e * (np.asarray(o.mask) != 0)
            if o.kind == 0:
                new = o.c0 * e + o.c1 * e + o.c2 * a["cur"] + o.c3 * a["prev"]
                a["prev"] = a["cur"].copy()
                a["cur"] = new
            else:
                a["cur"] = o.c0 * e + o.c1 * a["cur"]

    def set_tables(self, n_steps, amp=None, phasors=None):
        self.amp = None if amp is None else np.asarray(amp).reshape(n_steps, -1)
        self.ph = None if phasors is None else np.asarray(phasors).reshape(n_steps, -1)
        self.n_tab, self.cursor = n_steps, 0
        for m in self.mon:
            m["rec"] = []
        for f in self.fluxes:
            f["out"] = []

    def _sl(self, op):
        return tuple(slice(l, h) for l, h in zip(op.lo, op.hi))

    def run(self, n):
        if (self.src or self.mon) and self.cursor + n > self.n_tab:
            raise RuntimeError("no tabled steps left")
        Ca, Cb, Da, Db = self.coeffs
        for _ in range(n):
            kernels.step(self.F, (Ca, Cb, Da, Db), self.spacing + ((0.0,) if len(self.spacing) == 2 else ()), self.ndim == 2)
            s = self.cursor
            for g in sorted({o.group for o in self.src}):
                for o in (o for o in self.src if o.group == g):
                    a = self.amp[s, o.table]
                    if o.profile is not None:
                        a = a * np.asarray(o.profile)
                        if o.divisor != 1.0:
                            a = a / o.divisor
                    self.F[o.component][self._sl(o)] += a
            for m in self.mon:
                o = m["op"]
                d = self.F[o.component][self._sl(o)].copy()
                if o.record:
                    m["rec"].append(d)
                for k in range(o.n_freq):
                    ph = self.ph[s, o.phasor_col + k]
                    m["dft"][k] += (d * ph.real) * self.dt + 1j * ((d * ph.imag) * self.dt)
            self._run_flux()
            self._run_ade()
            self.cursor += 1
            self.kernel_launches += 2

    def update_h(self):
        kernels.update_h(self.F, self.coeffs[2], self.coeffs[3], self.spacing, self.ndim == 2)

    def update_e(self):
        kernels.update_e(self.F, self.coeffs[0], self.coeffs[1], self.spacing, self.ndim == 2)

    def sync(self):
        pass

    def records(self, i, steps):
        m = self.mon[i]
        return np.array(m["rec"][:steps]).reshape((steps,) + m["op"].shape)

    def dft(self, i):
        return self.mon[i]["dft"].copy()

    def set_dft(self, i, v):
        self.mon[i]["dft"][...] = v


class FakeSlabEngine(FakeEngine):
    """One x-slab of a 3-D grid on the CPU, with the ghost planes and the split sweep / post_step surface of
    the real engine, so SlabStepper's orchestration (halo exchange order, overlap split) can run under gloo.

    Non-last slabs keep an extended oracle domain of nx+3 planes: local planes, the neighbour's planes 0 and 1
    (ghosts), and a dummy; the oracle's own edge rules then only ever corrupt ghost planes, which the next
    exchange overwrites."""

    def __init__(self, ndim, dims, spacing, dt, dtype="float64", device=0, nx_global=None, x_offset=0, flags=0):
        assert ndim == 3
        self.nx_global = nx_global or dims[0]
        self.x_offset = x_offset
        self.last = x_offset + dims[0] == self.nx_global
        super().__init__(ndim, dims, spacing, dt, dtype)
        self.ext = dims[0] + (0 if self.last else 3)
        ed = (self.ext, dims[1], dims[2])
        self.G = {c: np.zeros(self._shape(ed, c)) for c in COMPONENTS}
        eps0, mu0 = 8.854187817e-12, 4 * np.pi * 1e-7
        self.coeffs = [np.full(ed, v) for v in (1.0, dt / eps0, 1.0, dt / mu0)]
        self.F = {c: self.G[c][: self.field_shape(c)[0]] for c in COMPONENTS}     # local views
        self.pending = []

    def set_uniform_coeffs(self, ca, cb, da, db):
        ed = (self.ext, self.dims[1], self.dims[2])
        self.coeffs = [np.full(ed, float(v)) for v in (ca, cb, da, db)]

    def _set_painted(self, planes):
        """The real engine holds nx (+1 with a right neighbour) coefficient planes; the extended oracle domain also
        computes throw-away ghost planes: pad by repeating the last plane supplied."""
        assert planes == self.dims[0] + (0 if self.last else 1)

        def pad(a):
            return np.concatenate([a] + [a[-1:]] * (self.ext - a.shape[0]), axis=0) if a.shape[0] < self.ext else a

        self.coeffs = [tuple(pad(c) for c in a) if isinstance(a, tuple) else pad(a) for a in self._painted]

    @staticmethod
    def _shape(d, comp):
        n = list(d)
        for ax in SHORT_AXES[comp]:
            n[ax] -= 1
        return tuple(n)

    def field_shape(self, comp):
        n = list(self.dims)
        for ax in SHORT_AXES[comp]:
            if ax == 0 and not self.last:
                continue
            n[ax] -= 1
        return tuple(n)

    def halo_tensors(self, spec):
        import torch

        nx = self.dims[0]
        return [(torch.from_numpy(self.G[c][0:p]), torch.from_numpy(self.G[c][nx:nx + p])) for c, p in spec]

    def sweep(self, i_begin, i_end, flip, stream=0):
        self.pending.append((i_begin, i_end))
        if not flip:
            return
        assert self.pending[0][0] == 0 and self.pending[-1][1] == self.dims[0]
        assert all(a[1] == b[0] for a, b in zip(self.pending, self.pending[1:]))
        self.pending = []
        kernels.step(self.G, self.coeffs, self.spacing, False)

    def post_step(self, stream=0):
        s = self.cursor
        for g in sorted({o.group for o in self.src}):
            for o in (o for o in self.src if o.group == g):
                if getattr(o, "ghost", False):
                    continue
                a = self.amp[s, o.table]
                if o.profile is not None:
                    a = a * np.asarray(o.profile)
                    if o.divisor != 1.0:
                        a = a / o.divisor
                self.F[o.component][self._sl(o)] += a
        for m in self.mon:
            o = m["op"]
            d = self.F[o.component][self._sl(o)].copy()
            if o.record:
                m["rec"].append(d)
            for k in range(o.n_freq):
                ph = self.ph[s, o.phasor_col + k]
                m["dft"][k] += (d * ph.real) * self.dt + 1j * ((d * ph.imag) * self.dt)
        self.cursor += 1


class FakeTensorLib:
    """Stands in for libfdtd_b200.so's fdtd_tensor_update on machines without a GPU: same argument protocol (ctypes
    pointer tables, mode bits), NumPy arithmetic in the kernel's operation order (csrc/fdtd_tensor.cuh)."""

    def __init__(self):
        self.calls = []

    def fdtd_last_error(self):
        return b""

    def fdtd_tensor_update(self, device, dtype, n, fp, cp, op, s, negative, mode, coef, arrs):
        import ctypes as C

        T = np.float32 if dtype == 0 else np.float64
        self.calls.append((dtype, n, negative, mode))

        def arr(ptr):
            return None if not ptr else np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float if T == np.float32 else C.c_double)), (n,))

        f, c, o = [arr(p) for p in fp], [arr(p) for p in cp], [arr(p) for p in op]
        ca = [arr(p) for p in arrs]
        s = T(s)
        full, mul32, div32 = mode & 1, (mode >> 1) & 1, (mode >> 2) & 1
        for k in range(3):
            if f[k] is None:
                continue
            if not full:
                d = ca[k] if ca[k] is not None else T(coef[k])
                if mul32 and T == np.float64:
                    t = (np.float32(s) * c[k].astype(np.float32)).astype(np.float64)
                else:
                    t = s * c[k]
                if div32 and T == np.float64:
                    t = (t.astype(np.float32) / np.asarray(d).astype(np.float32)).astype(np.float64)
                else:
                    t = t / d
                o[k][:] = f[k] - t if negative else f[k] + t
            else:
                r = [ca[3 * k + j] if ca[3 * k + j] is not None else T(coef[3 * k + j]) for j in range(3)]
                o[k][:] = f[k] + s * ((r[0] * c[0] + r[1] * c[1]) + r[2] * c[2])
        return 0
