"""Host-side Yee-grid bookkeeping, interface-compatible with the reference's ``GridSpec`` / ``YeeGrid``.

Mirrors the public surface of /root/reference/src/prismo/core/grid.py (class and method names,
argument meaning, quirks) because index boxes of sources and monitors are defined by it:
  * ``N = ceil(L * res)`` in floating point, ``N_total = N + 2*pml`` (grid.py:96-111)
  * 2-D <=> ``Lz == 0``: ``Nz = 1``, ``dz = 0`` (grid.py:92, 103-105)
  * ``point_to_index`` rounds, then clamps to the PHYSICAL count ``N-1`` (grid.py:345-352)
  * per-component boxes are clipped to the staggered array shape (grid.py:409-510)
This module is pure host arithmetic on a handful of integers; nothing here touches field data.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Union

import numpy as np

C0 = 299792458.0
COMPONENTS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
# array axes that are one element shorter than the grid, per component (grid.py:157-168)
SHORT_AXES = {"Ex": (1, 2), "Ey": (0, 2), "Ez": (0, 1), "Hx": (0,), "Hy": (1,), "Hz": (2,)}


@dataclass
class GridSpec:
    size: tuple
    resolution: Union[float, tuple]
    boundary_layers: int = 10

    def __post_init__(self):
        if any(s < 0 for s in self.size):
            raise ValueError("Grid size components must be non-negative")
        if isinstance(self.resolution, (int, float)):
            self.resolution = (self.resolution,) * 3
        elif len(self.resolution) != 3:
            raise ValueError("Resolution must be scalar or 3-tuple")
        if any(r <= 0 for r in self.resolution):
            raise ValueError("Resolution must be positive")


class YeeGrid:
    def __init__(self, spec: GridSpec):
        self.spec = spec
        self.Lx, self.Ly, self.Lz = spec.size
        self.res_x, self.res_y, self.res_z = spec.resolution
        self.dx, self.dy, self.dz = 1.0 / self.res_x, 1.0 / self.res_y, 1.0 / self.res_z
        self.is_2d = self.Lz == 0.0
        self.is_3d = not self.is_2d
        self.Nx = int(np.ceil(self.Lx * self.res_x)) if self.Lx > 0 else 1
        self.Ny = int(np.ceil(self.Ly * self.res_y)) if self.Ly > 0 else 1
        self.Nz = int(np.ceil(self.Lz * self.res_z)) if self.is_3d and self.Lz > 0 else 1
        if self.is_2d:
            self.dz = 0.0
        self.pml_layers = p = int(spec.boundary_layers)
        self.Nx_total = self.Nx + 2 * p
        self.Ny_total = self.Ny + 2 * p
        self.Nz_total = self.Nz + (2 * p if self.is_3d else 0)
        self.origin = np.array([-p * self.dx, -p * self.dy, -p * self.dz if self.is_3d else 0.0])

    @classmethod
    def like(cls, other) -> "YeeGrid":
        """Rebuild from any object exposing the reference YeeGrid's ``spec`` (duck-typed)."""
        if isinstance(other, cls):
            return other
        s = other.spec
        return cls(GridSpec(tuple(s.size), tuple(s.resolution) if not isinstance(s.resolution, (int, float))
                            else s.resolution, int(s.boundary_layers)))

    # ---- sizes -------------------------------------------------------------------------------
    @property
    def dimensions(self):
        return (self.Nx_total, self.Ny_total, self.Nz_total)

    @property
    def physical_dimensions(self):
        return (self.Nx, self.Ny, self.Nz)

    @property
    def spacing(self):
        return (self.dx, self.dy, self.dz)

    def get_field_shape(self, component):
        if component not in SHORT_AXES:
            raise ValueError(f"Unknown field component: {component}")
        n = list(self.dimensions)
        for ax in SHORT_AXES[component]:
            if ax < 2 or self.is_3d:
                n[ax] -= 1
        return tuple(n) if self.is_3d else tuple(n[:2])

    def get_coordinates(self, component):
        ax = [self.origin[d] + np.arange(n) * h for d, (n, h) in enumerate(zip(self.dimensions, self.spacing))]
        if component in ("Ey", "Ez", "Hy", "Hz"):
            ax[0] = ax[0][:-1] + self.dx / 2
        if component in ("Ex", "Ez", "Hx", "Hz"):
            ax[1] = ax[1][:-1] + self.dy / 2
        if component in ("Ex", "Ey", "Hx", "Hy") and self.is_3d:
            ax[2] = ax[2][:-1] + self.dz / 2
        return tuple(ax) if self.is_3d else tuple(ax[:2])

    # ---- time step -----------------------------------------------------------------------------
    def _inv_sq(self):
        s = (1 / self.dx) ** 2 + (1 / self.dy) ** 2
        return s if self.is_2d else s + (1 / self.dz) ** 2

    def get_courant_number(self, dt):
        return C0 * dt * np.sqrt(self._inv_sq())

    def suggest_time_step(self, safety_factor=0.9):
        return safety_factor / (C0 * np.sqrt(self._inv_sq()))

    get_time_step = suggest_time_step

    # ---- indices ---------------------------------------------------------------------------------
    def point_to_index(self, point):
        x, y, z = point
        i = int(round((x - self.origin[0]) / self.dx))
        j = int(round((y - self.origin[1]) / self.dy))
        k = int(round((z - self.origin[2]) / self.dz)) if self.is_3d else 0
        return (max(0, min(i, self.Nx - 1)), max(0, min(j, self.Ny - 1)), max(0, min(k, self.Nz - 1)))

    def index_to_coord(self, dim, indices):
        if dim not in (0, 1, 2):
            raise ValueError(f"Invalid dimension: {dim}. Must be 0, 1, or 2.")
        return self.origin[dim] + indices * self.spacing[dim]

    def region_bounds(self, center, size):
        """Half-open index bounds of a centre/size region (sources/base.py:97-139, monitors/base.py:93-134)."""
        lo = self.point_to_index(tuple(c - s / 2 for c, s in zip(center, size)))
        hi = self.point_to_index(tuple(c + s / 2 for c, s in zip(center, size)))
        x0, y0, z0 = lo
        x1, y1, z1 = hi
        if x0 == x1:
            x1 = x0 + 1
        if y0 == y1:
            y1 = y0 + 1
        if z0 == z1 and self.is_3d:
            z1 = z0 + 1
        return x0, x1, y0, y1, z0, z1

    def component_box(self, component, x_min, x_max, y_min, y_max, z_min, z_max):
        """The half-open box that ``get_component_indices`` enumerates: ((lo,hi), ...) per array axis."""
        if component not in SHORT_AXES:
            raise ValueError(f"Invalid field component: {component}")
        x_min = max(0, min(x_min, self.Nx - 1))
        x_max = max(1, min(x_max, self.Nx))
        y_min = max(0, min(y_min, self.Ny - 1))
        y_max = max(1, min(y_max, self.Ny))
        if self.is_3d:
            z_min = max(0, min(z_min, self.Nz - 1))
            z_max = max(1, min(z_max, self.Nz))
        else:
            z_min, z_max = 0, 1
        shape = self.get_field_shape(component)
        box = [[x_min, x_max], [y_min, y_max], [z_min, z_max]][: len(shape)]
        for ax in SHORT_AXES[component]:
            if ax < len(shape):
                box[ax][1] = min(box[ax][1], shape[ax])
        return tuple((a, max(a, b)) for a, b in box)

    def get_component_indices(self, component, x_min, x_max, y_min, y_max, z_min, z_max):
        return np.ix_(*[range(a, b) for a, b in self.component_box(component, x_min, x_max, y_min, y_max, z_min, z_max)])

    def is_inside_pml(self, i, j, k=0):
        p = self.pml_layers
        if i < p or i >= self.Nx_total - p or j < p or j >= self.Ny_total - p:
            return True
        return bool(self.is_3d and (k < p or k >= self.Nz_total - p))

    def get_physical_indices(self):
        p = self.pml_layers
        return (slice(p, self.Nx_total - p), slice(p, self.Ny_total - p),
                slice(p, self.Nz_total - p) if self.is_3d else slice(None))

    def __repr__(self):
        return (f"YeeGrid({'2D' if self.is_2d else '3D'}, shape={self.dimensions}, "
                f"spacing=({self.dx:.2e}, {self.dy:.2e}, {self.dz:.2e}), PML={self.pml_layers})")
