"""Geometry -> update coefficients ON THE DEVICE (first half).

The reference paints structures on the host (geometry/shapes.py): ``mask = shape.rasterize(x, y, z)`` builds the whole
(N, 3) fp64 point list, user code writes ``eps_rel[mask] = material.epsilon_r`` shape after shape, and ``MaxwellUpdater``
turns the painted arrays into Ca, Cb, Da, Db (core/solver.py:113-133).  Here the SHAPE LIST is what crosses the
boundary: ``fdtd_rasterize`` evaluates every shape at every cell's own coordinate on the GPU and writes the coefficient
arrays in the engine's layout — bit-identical masks and coefficients (tests/test_raster.py), no full-grid host array.

The classes below mirror the reference's constructors (same names, arguments, attributes and errors) so they can be
used where the reference package is not installed; ``lower_shapes`` is duck-typed by class name, so the reference's own
``prismo.geometry.shapes`` objects lower the same way.  There is no host rasteriser here: ``contains`` / ``rasterize``
of the mirrors raise — masks come from the device (``Engine.rasterize`` + ``download_coeffs``) or from the reference.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

KIND = {"Box": 0, "Sphere": 1, "Cylinder": 2, "Polygon": 3}
COMBINE = {"union": 1, "intersection": 2, "difference": 3}


@dataclass
class Material:
    """geometry/shapes.py:15-33.  epsilon_r may be a 3-sequence (eps_xx, eps_yy, eps_zz): a diagonal tensor, applied as
    per-component Cb inside the E stage (opt-in extension).  sigma_e / sigma_m extend the reference's dataclass with the
    conductivities MaxwellUpdater's material arrays carry (core/solver.py:84-110)."""

    name: str
    epsilon_r: object = 1.0
    mu_r: float = 1.0
    sigma_e: float = 0.0
    sigma_m: float = 0.0


class Shape:
    def __init__(self, material, center):
        self.material = material
        self.center = np.array(center)

    def contains(self, points):
        raise NotImplementedError("prismo_b200 rasterises on the device: use Engine.rasterize / Session.set_geometry")

    def rasterize(self, x, y, z=None):
        raise NotImplementedError("prismo_b200 rasterises on the device: use Engine.rasterize / Session.set_geometry")


class Box(Shape):
    def __init__(self, material, center, size):
        super().__init__(material, center)
        self.size = np.array(size)
        self.half_size = self.size / 2.0


class Sphere(Shape):
    def __init__(self, material, center, radius):
        super().__init__(material, center)
        self.radius = radius


class Cylinder(Shape):
    def __init__(self, material, center, radius, height, axis="z"):
        super().__init__(material, center)
        self.radius = radius
        self.height = height
        self.axis = axis.lower()
        if self.axis not in ["x", "y", "z"]:
            raise ValueError("axis must be 'x', 'y', or 'z'")


class Polygon(Shape):
    def __init__(self, material, vertices, z_min=-np.inf, z_max=np.inf):
        vertices = np.asarray(vertices)
        center = np.mean(vertices, axis=0)
        super().__init__(material, np.array([center[0], center[1], (z_min + z_max) / 2]))
        self.vertices = vertices
        self.z_min = z_min
        self.z_max = z_max


class GeometryGroup:
    """geometry/shapes.py:311-378.  ``material`` (not in the reference, whose groups only return masks) is what the
    combined mask is painted with; without it the group's first shape's material is used."""

    def __init__(self, shapes, operation="union", material=None):
        self.shapes = shapes
        self.operation = operation.lower()
        self.material = material
        if self.operation not in ["union", "intersection", "difference"]:
            raise ValueError("operation must be 'union', 'intersection', or 'difference'")


class ShapeStruct(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("axis", C.c_int32), ("combine", C.c_int32), ("paint", C.c_int32),
        ("center", C.c_double * 3), ("a", C.c_double * 3), ("vert_first", C.c_int32), ("vert_count", C.c_int32),
        ("eps_r", C.c_double * 3), ("mu_r", C.c_double), ("sigma_e", C.c_double), ("sigma_m", C.c_double),
    ]


def _material_values(m):
    e = getattr(m, "epsilon_r", 1.0)
    e3 = tuple(float(v) for v in np.asarray(e, dtype=np.float64).reshape(-1)) if np.ndim(e) > 0 else (float(e),) * 3
    if len(e3) != 3:
        raise ValueError("epsilon_r must be a scalar or (eps_xx, eps_yy, eps_zz)")
    return e3, float(getattr(m, "mu_r", 1.0)), float(getattr(m, "sigma_e", 0.0)), float(getattr(m, "sigma_m", 0.0))


def _primitive(shape, verts):
    name = next((k.__name__ for k in type(shape).__mro__ if k.__name__ in KIND), None)
    if name is None:
        raise TypeError(f"{type(shape).__name__} cannot be rasterised on the device (Box, Sphere, Cylinder, Polygon and "
                        "GeometryGroup can; CustomShape holds a host callable)")
    s = ShapeStruct()
    s.kind = KIND[name]
    c = np.asarray(shape.center, dtype=np.float64)
    s.center[:] = [float(c[0]), float(c[1]), float(c[2]) if c.size > 2 else 0.0]
    if name == "Box":
        h = np.asarray(shape.half_size, dtype=np.float64)                     # size / 2.0, formed once as in shapes.py:123
        s.a[:] = [float(h[0]), float(h[1]), float(h[2])]
    elif name == "Sphere":
        s.a[:] = [float(shape.radius), 0.0, 0.0]
    elif name == "Cylinder":
        s.axis = {"x": 0, "y": 1, "z": 2}[shape.axis]
        s.a[:] = [float(shape.radius), float(shape.height / 2), 0.0]       # shapes.py:214 compares with height / 2
    else:
        v = np.asarray(shape.vertices, dtype=np.float64)
        s.a[:] = [float(shape.z_min), float(shape.z_max), 0.0]
        s.vert_first, s.vert_count = len(verts), len(v)
        verts.extend((float(p[0]), float(p[1])) for p in v)
    return s


def lower_shapes(shapes: Sequence) -> tuple:
    """Shape list (painted in order; later shapes win) -> (ctypes array of fdtd_shape, vertex array (n, 2))."""
    out, verts = [], []
    for sh in shapes:
        if type(sh).__name__ == "GeometryGroup" or hasattr(sh, "operation"):
            members = list(sh.shapes)
            if not members:
                continue
            mat = getattr(sh, "material", None) or members[0].material
            for n, m in enumerate(members):
                s = _primitive(m, verts)
                s.combine = 0 if n == 0 else COMBINE[sh.operation]
                s.paint = int(n == len(members) - 1)
                if s.paint:
                    e3, s.mu_r, s.sigma_e, s.sigma_m = _material_values(mat)
                    s.eps_r[:] = e3
                out.append(s)
            continue
        s = _primitive(sh, verts)
        s.combine, s.paint = 0, 1
        e3, s.mu_r, s.sigma_e, s.sigma_m = _material_values(sh.material)
        s.eps_r[:] = e3
        out.append(s)
    arr = (ShapeStruct * max(len(out), 1))(*out)
    return arr, len(out), np.asarray(verts, dtype=np.float64).reshape(-1, 2)


def background_values(background) -> np.ndarray:
    """(eps_r | (exx, eyy, ezz), mu_r, sigma_e, sigma_m) or a Material -> the six doubles fdtd_rasterize takes."""
    if background is None:
        background = (1.0, 1.0, 0.0, 0.0)
    if hasattr(background, "epsilon_r"):
        e3, mu, se, sm = _material_values(background)
    else:
        b = tuple(background)
        if len(b) != 4:
            raise ValueError("background = (eps_r, mu_r, sigma_e, sigma_m)")
        e = b[0]
        e3 = tuple(float(v) for v in e) if np.ndim(e) > 0 else (float(e),) * 3
        mu, se, sm = float(b[1]), float(b[2]), float(b[3])
    return np.array([*e3, mu, se, sm], dtype=np.float64)


def cell_coordinates(grid) -> tuple:
    """Cell coordinates of the material arrays (shape grid.dimensions): origin + arange(N) * d, the base arrays of
    YeeGrid.get_coordinates (core/grid.py:194-197).  2-D grids return (x, y, None)."""
    d = grid.dimensions
    sp = grid.spacing
    ax = [grid.origin[n] + np.arange(d[n]) * sp[n] for n in range(3)]
    return ax[0], ax[1], (None if getattr(grid, "is_2d", False) else ax[2])
