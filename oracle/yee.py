"""Oracle for the OPT-IN physics mode (stable Yee leap-frog + CPML).  Test infrastructure.  PARITY UNPINNED:
this is our own CPU restatement of prismo_b200/csrc/fdtd_yee.cuh — the reference has no executed counterpart
(its update is unstable; its CPML is a stub, F6) — so it is validated by physics checks, not goldens.

Scheme on the reference's staggering (core/grid.py:60-66): E update = the reference's forward differences
(core/solver.py:255-309); H update = BACKWARD differences; every derivative d becomes ki*d + psi with
psi <- b*psi + a*d, (b, a, ki) = per-axis vectors that are the identity (0, 0, 1) outside the layers, so full-size
psi arrays give exactly what the device computes on slabs only.
"""
from __future__ import annotations

import numpy as np


def _bc(v, axis):
    s = [1, 1, 1]
    s[axis] = -1
    return v.reshape(s)


class YeeOracle:
    def __init__(self, dims, spacing, dt, coeffs, cpml_axes=None):
        """coeffs = (Ca, Cb, Da, Db) uniform scalars; cpml_axes = [(6,N) arrays for x, y, z] or None."""
        self.dims, self.sp, self.dt = dims, spacing, dt
        self.ca, self.cb, self.da, self.db = coeffs
        nx, ny, nz = dims
        shp = {"Ex": (nx, ny - 1, nz - 1), "Ey": (nx - 1, ny, nz - 1), "Ez": (nx - 1, ny - 1, nz),
               "Hx": (nx - 1, ny, nz), "Hy": (nx, ny - 1, nz), "Hz": (nx, ny, nz - 1)}
        self.F = {c: np.zeros(s) for c, s in shp.items()}
        if cpml_axes is None:
            cpml_axes = [np.stack([np.zeros(n), np.zeros(n), np.ones(n)] * 2) for n in dims]
        self.cx = cpml_axes
        self.psi = {}

    def _d(self, key, d, axis, pos, lo):
        """CPML-transform derivative array d whose index along `axis` starts at grid index lo; pos 0 = E rows."""
        b, a, ki = (self.cx[axis][pos + q][lo:lo + d.shape[axis]] for q in range(3))
        p = self.psi.get(key)
        if p is None:
            p = self.psi[key] = np.zeros_like(d)
        p[...] = _bc(b, axis) * p + _bc(a, axis) * d
        return _bc(ki, axis) * d + p

    def step(self):
        F, (dx, dy, dz) = self.F, self.sp
        Ex, Ey, Ez, Hx, Hy, Hz = (F[c] for c in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"))
        # ---- H: backward differences, interior rows only -------------------------------------------------
        d1 = self._d("hx_y", (Ez[:, 1:, 1:-1] - Ez[:, :-1, 1:-1]) / dy, 1, 3, 1)        # j = 1..ny-2, k = 1..nz-2
        d2 = self._d("hx_z", (Ey[:, 1:-1, 1:] - Ey[:, 1:-1, :-1]) / dz, 2, 3, 1)
        Hx[:, 1:-1, 1:-1] = self.da * Hx[:, 1:-1, 1:-1] - self.db * (d1 - d2)
        d1 = self._d("hy_z", (Ex[1:-1, :, 1:] - Ex[1:-1, :, :-1]) / dz, 2, 3, 1)        # i = 1..nx-2, k = 1..nz-2
        d2 = self._d("hy_x", (Ez[1:, :, 1:-1] - Ez[:-1, :, 1:-1]) / dx, 0, 3, 1)
        Hy[1:-1, :, 1:-1] = self.da * Hy[1:-1, :, 1:-1] - self.db * (d1 - d2)
        d1 = self._d("hz_x", (Ey[1:, 1:-1, :] - Ey[:-1, 1:-1, :]) / dx, 0, 3, 1)        # i = 1..nx-2, j = 1..ny-2
        d2 = self._d("hz_y", (Ex[1:-1, 1:, :] - Ex[1:-1, :-1, :]) / dy, 1, 3, 1)
        Hz[1:-1, 1:-1, :] = self.da * Hz[1:-1, 1:-1, :] - self.db * (d1 - d2)
        # ---- E: forward differences, whole arrays ------------------------------------------------------------
        d1 = self._d("ex_y", (Hz[:, 1:, :] - Hz[:, :-1, :]) / dy, 1, 0, 0)
        d2 = self._d("ex_z", (Hy[:, :, 1:] - Hy[:, :, :-1]) / dz, 2, 0, 0)
        Ex[...] = self.ca * Ex + self.cb * (d1 - d2)
        d1 = self._d("ey_z", (Hx[:, :, 1:] - Hx[:, :, :-1]) / dz, 2, 0, 0)
        d2 = self._d("ey_x", (Hz[1:, :, :] - Hz[:-1, :, :]) / dx, 0, 0, 0)
        Ey[...] = self.ca * Ey + self.cb * (d1 - d2)
        d1 = self._d("ez_x", (Hy[1:, :, :] - Hy[:-1, :, :]) / dx, 0, 0, 0)
        d2 = self._d("ez_y", (Hx[:, 1:, :] - Hx[:, :-1, :]) / dy, 1, 0, 0)
        Ez[...] = self.ca * Ez + self.cb * (d1 - d2)
