"""Oracle restatement of the reference's geometry rasterisation.  Test infrastructure only.

Follows /root/reference/src/prismo/geometry/shapes.py:
  Shape.rasterize          :71-99     meshgrid(indexing="ij") -> (N, 3) points (z = 0 in 2-D) -> contains() -> reshape
  Box.contains             :125-132   all(|p - c| <= size / 2)
  Sphere.contains          :155-159   np.linalg.norm(p - c, axis=1) <= radius
  Cylinder.contains        :194-214   sqrt(da**2 + db**2) <= radius  &  |p_ax - c_ax| <= height / 2
  Polygon.contains         :249-283   ray casting in xy, z_min <= z <= z_max
  GeometryGroup.rasterize  :338-378   union / intersection / difference of the member masks
followed by what the reference's users do with the masks (eps_rel[mask] = material.epsilon_r, in list order) and by
MaxwellUpdater._compute_update_coefficients (core/solver.py:113-133, restated in oracle/kernels.py::coefficients).

Unlike the reference it never materialises the (N, 3) point list: each test is evaluated on broadcast coordinate
axes with the SAME elementwise fp64 operations in the same order, so the masks are bit-identical
(tests/test_raster.py pins that against the live reference, adversarial boundary cells included).

A shape is a dict: kind in {"box", "sphere", "cylinder", "polygon", "group"} + its parameters + material values
(eps_r scalar or 3-tuple, mu_r, sigma_e, sigma_m).
"""
from __future__ import annotations

import numpy as np

from . import kernels


def _axes(x, y, z):
    x = np.asarray(x, dtype=np.float64)[:, None, None]
    y = np.asarray(y, dtype=np.float64)[None, :, None]
    z = (np.zeros(1) if z is None else np.asarray(z, dtype=np.float64))[None, None, :]
    return x, y, z


def contains(shape, x, y, z):
    """Boolean mask (len(x), len(y), len(z) or 1) of one primitive or group."""
    X, Y, Z = _axes(x, y, z)
    full = np.broadcast_shapes(X.shape, Y.shape, Z.shape)
    kind = shape["kind"]
    if kind == "group":                                                     # shapes.py:353-376
        masks = [contains(s, x, y, z) for s in shape["shapes"]]
        out = masks[0]
        for m in masks[1:]:
            out = {"union": out | m, "intersection": out & m, "difference": out & ~m}[shape["operation"]]
        return out
    if kind == "polygon":                                                   # shapes.py:249-283
        v = np.asarray(shape["vertices"], dtype=np.float64)
        zin = (Z >= shape["z_min"]) & (Z <= shape["z_max"])
        count = np.zeros(np.broadcast_shapes(X.shape, Y.shape), dtype=np.int64)
        n = len(v)
        with np.errstate(divide="ignore", invalid="ignore"):
            for q in range(n):
                v1, v2 = v[q], v[(q + 1) % n]
                straddle = (v1[1] > Y) != (v2[1] > Y)
                xc = (v2[0] - v1[0]) * (Y - v1[1]) / (v2[1] - v1[1]) + v1[0]
                count += straddle & (X < xc)
        return np.broadcast_to((count % 2 == 1) & zin, full).copy()
    c = np.asarray(shape["center"], dtype=np.float64)
    dx, dy, dz = X - c[0], Y - c[1], Z - c[2]
    if kind == "box":                                                       # shapes.py:125-132
        h = np.asarray(shape["size"], dtype=np.float64) / 2.0
        return np.broadcast_to((np.abs(dx) <= h[0]) & (np.abs(dy) <= h[1]) & (np.abs(dz) <= h[2]), full).copy()
    if kind == "sphere":                                                    # shapes.py:155-159: add.reduce over a row's 3 squares
        return np.broadcast_to(np.sqrt((dx * dx + dy * dy) + dz * dz) <= shape["radius"], full).copy()
    if kind == "cylinder":                                                  # shapes.py:194-214
        ax = {"x": 0, "y": 1, "z": 2}[shape.get("axis", "z").lower()]
        d = (dx, dy, dz)
        p = [i for i in range(3) if i != ax]
        radial = np.sqrt(d[p[0]] ** 2 + d[p[1]] ** 2)
        return np.broadcast_to((radial <= shape["radius"]) & (np.abs(d[ax]) <= shape["height"] / 2), full).copy()
    raise ValueError(kind)


def material_arrays(shapes, x, y, z, background=(1.0, 1.0, 0.0, 0.0)):
    """eps_rel (a 3-tuple of arrays when any material is anisotropic), mu_rel, sigma_e, sigma_m painted in list order."""
    nz = 1 if z is None else len(z)
    dims = (len(x), len(y), nz)
    bg_eps = background[0]
    aniso = np.ndim(bg_eps) > 0 or any(np.ndim(s.get("eps_r", 1.0)) > 0 for s in shapes)
    bge = tuple(bg_eps) if np.ndim(bg_eps) > 0 else (bg_eps,) * 3
    eps = [np.full(dims, bge[d], dtype=np.float64) for d in range(3 if aniso else 1)]
    mu = np.full(dims, background[1], dtype=np.float64)
    se = np.full(dims, background[2], dtype=np.float64)
    sm = np.full(dims, background[3], dtype=np.float64)
    for s in shapes:
        m = contains(s, x, y, z)
        e = s.get("eps_r", 1.0)
        e3 = tuple(e) if np.ndim(e) > 0 else (e,) * 3
        for d in range(len(eps)):
            eps[d][m] = e3[d]
        mu[m] = s.get("mu_r", 1.0)
        se[m] = s.get("sigma_e", 0.0)
        sm[m] = s.get("sigma_m", 0.0)
    return (tuple(eps) if aniso else eps[0]), mu, se, sm


def coefficient_arrays(shapes, x, y, z, dt, background=(1.0, 1.0, 0.0, 0.0)):
    """Ca, Cb (a 3-tuple when anisotropic; Ca is then the eps_xx one), Da, Db — solver.py:113-133 on the painted arrays."""
    eps, mu, se, sm = material_arrays(shapes, x, y, z, background)
    if isinstance(eps, tuple):
        Ca, Cbx, Da, Db = kernels.coefficients(eps[0], mu, se, sm, dt)
        Cby = kernels.coefficients(eps[1], mu, se, sm, dt)[1]
        Cbz = kernels.coefficients(eps[2], mu, se, sm, dt)[1]
        return Ca, (Cbx, Cby, Cbz), Da, Db
    return kernels.coefficients(eps, mu, se, sm, dt)
