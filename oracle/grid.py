"""Oracle restatement of the reference Yee-grid bookkeeping (host side).  Test infrastructure.

Follows /root/reference/src/prismo/core/grid.py:
  sizes :78-120, shapes :137-176, dt :304-326, point_to_index :328-354,
  index_to_coord :356-381, get_component_indices :383-513,
and sources/base.py:97-139 == monitors/base.py:93-134 (region boxes).
"""
from __future__ import annotations

import math

import numpy as np

C0 = 299792458.0
COMPONENTS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
# axes along which a component's array is one shorter than the grid (grid.py:157-168)
_SHORT = {"Ex": (1, 2), "Ey": (0, 2), "Ez": (0, 1), "Hx": (0,), "Hy": (1,), "Hz": (2,)}


class OGrid:
    """size in metres, resolution in points/metre, pml = boundary layers per side."""

    def __init__(self, size, resolution, pml=10):
        if isinstance(resolution, (int, float)):
            resolution = (resolution,) * 3
        self.size = tuple(size)
        self.res = tuple(resolution)
        self.pml = int(pml)
        L, r = self.size, self.res
        self.is_2d = L[2] == 0.0
        self.is_3d = not self.is_2d
        self.dx, self.dy, self.dz = (1.0 / r[0], 1.0 / r[1], 1.0 / r[2])
        # float ceil, exactly as grid.py:96-100 (10e-6*50e6 -> 501)
        self.Nx = int(np.ceil(L[0] * r[0])) if L[0] > 0 else 1
        self.Ny = int(np.ceil(L[1] * r[1])) if L[1] > 0 else 1
        self.Nz = int(np.ceil(L[2] * r[2])) if (L[2] > 0 and not self.is_2d) else 1
        if self.is_2d:
            self.Nz = 1
            self.dz = 0.0
        p = self.pml
        self.nx = self.Nx + 2 * p
        self.ny = self.Ny + 2 * p
        self.nz = self.Nz + (2 * p if self.is_3d else 0)
        self.origin = np.array([-p * self.dx, -p * self.dy, -p * self.dz if self.is_3d else 0.0])

    # ---- shapes -----------------------------------------------------------------
    @property
    def dims(self):
        return (self.nx, self.ny, self.nz)

    @property
    def spacing(self):
        return (self.dx, self.dy, self.dz)

    def shape(self, comp):
        full = [self.nx, self.ny, self.nz]
        for ax in _SHORT[comp]:
            full[ax] -= 1
        if self.is_2d:
            # 2-D drops the z axis with no "-1" on it (grid.py:158-168)
            full = [self.nx, self.ny]
            for ax in _SHORT[comp]:
                if ax < 2:
                    full[ax] -= 1
        return tuple(full)

    # ---- time step --------------------------------------------------------------
    def time_step(self, safety=0.9):
        if self.is_2d:
            s = (1 / self.dx) ** 2 + (1 / self.dy) ** 2
        else:
            s = (1 / self.dx) ** 2 + (1 / self.dy) ** 2 + (1 / self.dz) ** 2
        return safety / (C0 * np.sqrt(s))

    # ---- index boxes ------------------------------------------------------------
    def point_to_index(self, p):
        x, y, z = p
        i = int(round((x - self.origin[0]) / self.dx))
        j = int(round((y - self.origin[1]) / self.dy))
        k = int(round((z - self.origin[2]) / self.dz)) if self.is_3d else 0
        # clamp to the PHYSICAL counts, not the totals (grid.py:350-352)
        i = max(0, min(i, self.Nx - 1))
        j = max(0, min(j, self.Ny - 1))
        k = max(0, min(k, self.Nz - 1))
        return i, j, k

    def index_to_coord(self, dim, idx):
        return self.origin[dim] + idx * (self.dx, self.dy, self.dz)[dim]

    def region(self, center, size):
        """(x0,x1,y0,y1,z0,z1) half-open, as sources/base.py:97-139."""
        h = [s / 2 for s in size]
        lo = self.point_to_index(tuple(c - d for c, d in zip(center, h)))
        hi = self.point_to_index(tuple(c + d for c, d in zip(center, h)))
        x0, y0, z0 = lo
        x1, y1, z1 = hi
        if x0 == x1:
            x1 = x0 + 1
        if y0 == y1:
            y1 = y0 + 1
        if z0 == z1 and self.is_3d:
            z1 = z0 + 1
        return x0, x1, y0, y1, z0, z1

    def component_box(self, comp, region):
        """Per-component clip of a region to half-open ranges (grid.py:409-510).

        Returns ((lo,hi),...) one pair per array axis (2 pairs in 2-D).  A range may be empty.
        """
        x0, x1, y0, y1, z0, z1 = region
        x0 = max(0, min(x0, self.Nx - 1))
        x1 = max(1, min(x1, self.Nx))
        y0 = max(0, min(y0, self.Ny - 1))
        y1 = max(1, min(y1, self.Ny))
        if self.is_3d:
            z0 = max(0, min(z0, self.Nz - 1))
            z1 = max(1, min(z1, self.Nz))
        else:
            z0, z1 = 0, 1
        shp = self.shape(comp)
        rng = [[x0, x1], [y0, y1], [z0, z1]]
        for ax in _SHORT[comp]:
            if ax < len(shp):
                rng[ax][1] = min(rng[ax][1], shp[ax])
        n = 3 if self.is_3d else 2
        return tuple((a, max(a, b)) for a, b in rng[:n])

    @staticmethod
    def box_slices(box):
        return tuple(slice(a, b) for a, b in box)
