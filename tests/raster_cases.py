"""Shared scenes for the rasterisation tests (oracle dicts <-> shape objects of a given module)."""
import numpy as np

# name -> (grid dims, spacing, origin, shape list as oracle dicts, background)
# Several parameters are chosen so that cells sit EXACTLY on a boundary (<= vs <, summation order of the sphere's norm).


def coords(dims, spacing, origin):
    ax = [origin[d] + np.arange(dims[d]) * spacing[d] for d in range(3)]
    return ax[0], ax[1], (ax[2] if dims[2] > 1 else None)


def _scene3d():
    dims, sp, org = (23, 19, 17), (0.05, 0.04, 0.06), (-0.5, -0.35, -0.45)
    x, y, z = coords(dims, sp, org)
    # sphere radius = the exact distance of cell (15, 9, 10) from the centre as np.linalg.norm computes it
    c = np.array([0.1, 0.02, 0.03])
    r = float(np.linalg.norm(np.array([[x[15], y[9], z[10]]]) - c, axis=1)[0])
    cyl_c = np.array([-0.2, 0.1, 0.0])
    rc = float(np.sqrt((x[3] - cyl_c[0]) ** 2 + (y[14] - cyl_c[1]) ** 2))
    shapes = [
        dict(kind="box", center=(0.0, 0.0, 0.0), size=(0.6, 2 * (y[12] - 0.0), 0.5), eps_r=2.25),
        dict(kind="sphere", center=tuple(c), radius=r, eps_r=11.9, sigma_e=30.0),
        dict(kind="cylinder", center=tuple(cyl_c), radius=rc, height=0.5, axis="z", eps_r=4.0, mu_r=1.5, sigma_m=2.0),
        dict(kind="cylinder", center=(0.2, -0.1, 0.1), radius=0.13, height=0.31, axis="x", eps_r=6.0),
        dict(kind="cylinder", center=(0.0, 0.0, -0.2), radius=0.09, height=2 * (y[13] - 0.0), axis="y", eps_r=3.0),
        dict(kind="polygon", vertices=[(-0.4, -0.3), (-0.1, -0.3), (-0.05, -0.1), (-0.25, 0.05), (-0.45, -0.12)],
             z_min=float(z[2]), z_max=0.2, eps_r=7.5),
        dict(kind="group", operation="difference", eps_r=9.0, shapes=[
            dict(kind="box", center=(0.3, 0.2, 0.2), size=(0.3, 0.3, 0.4)),
            dict(kind="sphere", center=(0.3, 0.2, 0.2), radius=0.12)]),
        dict(kind="group", operation="intersection", eps_r=5.0, sigma_e=1.0, shapes=[
            dict(kind="sphere", center=(-0.3, 0.2, -0.2), radius=0.2),
            dict(kind="box", center=(-0.3, 0.25, -0.2), size=(0.5, 0.2, 0.2))]),
        dict(kind="group", operation="union", eps_r=1.7, shapes=[
            dict(kind="box", center=(0.35, -0.25, -0.3), size=(0.1, 0.1, 0.1)),
            dict(kind="cylinder", center=(0.35, -0.25, -0.3), radius=0.08, height=0.3, axis="z")]),
    ]
    return dims, sp, org, shapes, (1.44, 1.0, 0.0, 0.0)


def _scene2d():
    dims, sp, org = (41, 37, 1), (0.025, 0.03, 0.0), (-0.5, -0.55, 0.0)
    x, y, _ = coords(dims, sp, org)
    shapes = [
        dict(kind="box", center=(0.0, 0.0, 0.0), size=(0.5, 0.42, 0.0), eps_r=2.0),          # Lz = 0: only z == 0 passes
        dict(kind="sphere", center=(0.1, -0.1, 0.0), radius=float(np.hypot(x[30] - 0.1, y[20] + 0.1)), eps_r=12.0),
        dict(kind="cylinder", center=(-0.2, 0.2, 0.0), radius=0.11, height=1.0, axis="z", eps_r=3.5, sigma_e=4.0),
        dict(kind="polygon", vertices=[(0.1, 0.1), (0.4, 0.15), (0.3, 0.4)], z_min=-1.0, z_max=1.0, eps_r=6.0),
    ]
    return dims, sp, org, shapes, (1.0, 1.0, 0.0, 0.0)


def _scene_aniso():
    dims, sp, org = (21, 18, 35), (0.05, 0.05, 0.05), (-0.5, -0.45, -0.85)
    shapes = [
        dict(kind="box", center=(0.0, 0.0, 0.0), size=(2.0, 0.5, 1.0), eps_r=(2.2, 2.31, 2.2)),      # crosses every x plane
        dict(kind="cylinder", center=(0.0, 0.0, 0.0), radius=0.2, height=0.8, axis="x", eps_r=12.1),
        dict(kind="sphere", center=(0.2, 0.1, -0.3), radius=0.15, eps_r=(4.0, 5.0, 6.0), mu_r=1.2, sigma_m=3.0),
    ]
    return dims, sp, org, shapes, (1.0, 1.0, 0.0, 0.0)


SCENES = {"scene3d": _scene3d(), "scene2d": _scene2d(), "aniso": _scene_aniso()}


def build_shapes(shapes, mod, material_cls=None):
    """Oracle dicts -> shape objects of `mod` (prismo_b200.geometry, or the reference's prismo.geometry.shapes)."""
    Material = material_cls or mod.Material

    def mat(s):
        e = s.get("eps_r", 1.0)
        m = Material("m", epsilon_r=e, mu_r=s.get("mu_r", 1.0))
        for k in ("sigma_e", "sigma_m"):                 # the reference's Material has no conductivities: set as attributes
            setattr(m, k, s.get(k, 0.0))
        return m

    def prim(s, m):
        k = s["kind"]
        if k == "box":
            return mod.Box(m, s["center"], s["size"])
        if k == "sphere":
            return mod.Sphere(m, s["center"], s["radius"])
        if k == "cylinder":
            return mod.Cylinder(m, s["center"], s["radius"], s["height"], s.get("axis", "z"))
        if k == "polygon":
            return mod.Polygon(m, np.array(s["vertices"]), s["z_min"], s["z_max"])
        raise ValueError(k)

    out = []
    for s in shapes:
        if s["kind"] == "group":
            g = mod.GeometryGroup([prim(q, mat(s)) for q in s["shapes"]], s["operation"])
            g.material = mat(s)
            out.append(g)
        else:
            out.append(prim(s, mat(s)))
    return out


DT = 1.0e-11        # Courant 0.1-0.2 on the scenes' spacings


def reference_scene(name, prismo):
    """Masks (one per list entry) and Ca, Cb, Da, Db of a scene from the REAL reference: Shape.rasterize /
    GeometryGroup.rasterize (geometry/shapes.py), painting in list order, MaxwellUpdater (core/solver.py:41-133)."""
    from prismo.core.grid import GridSpec, YeeGrid
    from prismo.core.solver import MaxwellUpdater
    from prismo.geometry import shapes as RS

    dims, sp, org, shapes, bg = SCENES[name]
    x, y, z = coords(dims, sp, org)
    eps = np.full(dims, bg[0], dtype=np.float64)
    mu = np.full(dims, bg[1], dtype=np.float64)
    se = np.full(dims, bg[2], dtype=np.float64)
    sm = np.full(dims, bg[3], dtype=np.float64)
    masks = []
    for obj in build_shapes(shapes, RS):
        m = obj.rasterize(x, y, z)
        m = m[0] if isinstance(m, tuple) else m          # GeometryGroup returns (combined, member masks)
        m = m.reshape(dims)
        masks.append(m)
        eps[m] = obj.material.epsilon_r
        mu[m] = obj.material.mu_r
        se[m] = obj.material.sigma_e
        sm[m] = obj.material.sigma_m
    res = tuple(1.0 / s if s > 0 else 1.0 for s in sp)
    size = tuple((dims[d] - 0.5) / res[d] if dims[d] > 1 else 0.0 for d in range(3))
    grid = YeeGrid(GridSpec(size=size, resolution=res, boundary_layers=0))
    assert tuple(grid.dimensions) == tuple(dims), (grid.dimensions, dims)
    upd = MaxwellUpdater(grid, DT, material_arrays=dict(eps_rel=eps, mu_rel=mu, sigma_e=se, sigma_m=sm), backend="numpy")
    return masks, tuple(np.asarray(a) for a in (upd.Ca, upd.Cb, upd.Da, upd.Db))
