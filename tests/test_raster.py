"""Row f4 (first half): geometry rasterisation on the device — shape list -> Ca, Cb, Da, Db.

CPU: the NumPy restatement (oracle/raster.py) is pinned against the LIVE reference (geometry/shapes.py masks,
MaxwellUpdater coefficients) and against the committed goldens of the real reference; host lowering of shape objects.
GPU: fdtd_rasterize is bit-exact against the oracle and the goldens (fp64; fp32 = the rounded fp64 values), a run on
device-painted coefficients equals a run on host-painted ones, and per-component Cb (opt-in) equals its oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import kernels, raster
from tests import raster_cases as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "raster.npz")


def _gold_masks(g, name, dims):
    n = int(g[f"{name}.n_masks"])
    bits = np.unpackbits(g[f"{name}.masks"])[: n * int(np.prod(dims))]
    return bits.reshape((n,) + tuple(dims)).astype(bool)


# ---- oracle pinned ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["scene3d", "scene2d"])
def test_oracle_raster_matches_golden(name):
    g = np.load(GOLD)
    dims, sp, org, shapes, bg = R.SCENES[name]
    x, y, z = R.coords(dims, sp, org)
    masks = _gold_masks(g, name, dims)
    for n, s in enumerate(shapes):
        assert np.array_equal(raster.contains(s, x, y, z), masks[n]), (name, n, s["kind"])
        assert 0 < masks[n].sum() < masks[n].size
    got = raster.coefficient_arrays(shapes, x, y, z, R.DT, bg)
    for k, a in zip(("Ca", "Cb", "Da", "Db"), got):
        assert np.array_equal(a, g[f"{name}.{k}"]), (name, k)


@pytest.mark.reference
@pytest.mark.parametrize("name", ["scene3d", "scene2d"])
def test_oracle_raster_matches_live_reference_and_golden_is_current(name, ref):
    g = np.load(GOLD)
    dims, sp, org, shapes, bg = R.SCENES[name]
    x, y, z = R.coords(dims, sp, org)
    masks, coefs = R.reference_scene(name, ref)
    gm = _gold_masks(g, name, dims)
    for n, s in enumerate(shapes):
        assert np.array_equal(raster.contains(s, x, y, z), masks[n]), (name, n)
        assert np.array_equal(gm[n], masks[n])
    for k, a, b in zip(("Ca", "Cb", "Da", "Db"), raster.coefficient_arrays(shapes, x, y, z, R.DT, bg), coefs):
        assert np.array_equal(a, b) and np.array_equal(b, g[f"{name}.{k}"]), (name, k)


@pytest.mark.reference
def test_oracle_sphere_and_box_boundaries_against_live_reference(ref):
    """Adversarial boundaries: a radius equal to the reference's own distance of many different cells (<= must hold
    for exactly the same cells: pins the summation order of np.linalg.norm), box faces exactly on cell coordinates."""
    from prismo.geometry import shapes as RS

    rng = np.random.default_rng(3)
    x = np.linspace(-0.37, 0.41, 14)
    y = np.linspace(-0.29, 0.33, 13)
    z = np.linspace(-0.45, 0.22, 12)
    mat = RS.Material("m", 2.0)
    for _ in range(40):
        c = rng.uniform(-0.1, 0.1, 3)
        i, j, k = rng.integers(0, 12, 3)
        r = float(np.linalg.norm(np.array([[x[i], y[j], z[k]]]) - c, axis=1)[0])
        want = RS.Sphere(mat, tuple(c), r).rasterize(x, y, z)
        got = raster.contains(dict(kind="sphere", center=tuple(c), radius=r), x, y, z)
        assert want[i, j, k] and np.array_equal(got, want)
        size = (2 * abs(x[i] - c[0]), 2 * abs(y[j] - c[1]), 2 * abs(z[k] - c[2]))
        want = RS.Box(mat, tuple(c), size).rasterize(x, y, z)
        got = raster.contains(dict(kind="box", center=tuple(c), size=size), x, y, z)
        assert np.array_equal(got, want)
        ax = "xyz"[int(rng.integers(0, 3))]
        want = RS.Cylinder(mat, tuple(c), r * 0.8, 0.4, ax).rasterize(x, y, z)
        got = raster.contains(dict(kind="cylinder", center=tuple(c), radius=r * 0.8, height=0.4, axis=ax), x, y, z)
        assert np.array_equal(got, want)


# ---- host lowering ---------------------------------------------------------------------------------------------------
def test_lowering_of_mirror_shapes():
    from prismo_b200 import geometry as G

    dims, sp, org, shapes, bg = R.SCENES["scene3d"]
    arr, n, verts = G.lower_shapes(R.build_shapes(shapes, G))
    assert n == 6 + 3 * 2 and verts.shape == (5, 2)
    kinds = [arr[q].kind for q in range(n)]
    assert kinds == [0, 1, 2, 2, 2, 3, 0, 1, 1, 0, 0, 2]
    assert [arr[q].combine for q in range(n)] == [0] * 6 + [0, 3, 0, 2, 0, 1]
    assert [arr[q].paint for q in range(n)] == [1] * 6 + [0, 1, 0, 1, 0, 1]
    assert arr[0].a[1] == shapes[0]["size"][1] / 2.0 and arr[1].a[0] == shapes[1]["radius"]
    assert arr[3].axis == 0 and arr[4].axis == 1 and arr[2].axis == 2 and arr[2].a[1] == 0.25
    assert arr[5].vert_first == 0 and arr[5].vert_count == 5 and arr[5].a[0] == shapes[5]["z_min"]
    assert arr[1].sigma_e == 30.0 and arr[2].mu_r == 1.5 and arr[2].sigma_m == 2.0 and tuple(arr[7].eps_r) == (9.0,) * 3
    assert tuple(G.background_values(bg)) == (1.44, 1.44, 1.44, 1.0, 0.0, 0.0)
    assert tuple(G.background_values(G.Material("bg", (2.0, 3.0, 4.0)))) == (2.0, 3.0, 4.0, 1.0, 0.0, 0.0)
    a2, n2, _ = G.lower_shapes(R.build_shapes(R.SCENES["aniso"][3], G))
    assert tuple(a2[0].eps_r) == (2.2, 2.31, 2.2) and tuple(a2[2].eps_r) == (4.0, 5.0, 6.0)
    with pytest.raises(ValueError):
        G.Cylinder(G.Material("m"), (0, 0, 0), 1.0, 1.0, axis="w")
    with pytest.raises(ValueError):
        G.GeometryGroup([], "xor")
    with pytest.raises(NotImplementedError):            # no host rasteriser in the product
        G.Box(G.Material("m"), (0, 0, 0), (1, 1, 1)).rasterize(np.zeros(2), np.zeros(2))

    class CustomShape(G.Shape):
        pass

    with pytest.raises(TypeError):
        G.lower_shapes([CustomShape(G.Material("m"), (0, 0, 0))])


@pytest.mark.reference
def test_lowering_accepts_the_reference_shape_objects(ref):
    from prismo.geometry import shapes as RS

    from prismo_b200 import geometry as G

    shapes = R.SCENES["scene3d"][3]
    a, n, v = G.lower_shapes(R.build_shapes(shapes, G))
    b, m, w = G.lower_shapes(R.build_shapes(shapes, RS))
    assert n == m and np.array_equal(v, w)
    assert bytes(a) == bytes(b)


def test_rasterize_argument_checks_need_no_gpu():
    from prismo_b200 import _lib

    lib = _lib.load()
    assert lib.fdtd_rasterize(None, None, 0, None, 0, None, None, None, 0, None) == -1
    assert lib.fdtd_download_coeffs(None, 0, None, 0) == -1
    assert lib.fdtd_set_coeffs_aniso(None, None, None, None, None, None, None, 0) == -1
    assert lib.fdtd_struct_size(4) == C.sizeof(__import__("prismo_b200.geometry", fromlist=["ShapeStruct"]).ShapeStruct)


# ---- device ----------------------------------------------------------------------------------------------------------
def _engine(pb, name, dtype, **kw):
    dims, sp, org, shapes, bg = R.SCENES[name]
    ndim = 3 if dims[2] > 1 else 2
    spacing = sp if ndim == 3 else (sp[0], sp[1], 1.0)
    return pb.Engine(ndim, dims, spacing, R.DT, dtype=dtype, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["scene3d", "scene2d"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_device_raster_matches_oracle_and_reference_golden(name, dtype):
    import prismo_b200 as pb
    from prismo_b200 import geometry as G

    g = np.load(GOLD)
    dims, sp, org, shapes, bg = R.SCENES[name]
    x, y, z = R.coords(dims, sp, org)
    want = raster.coefficient_arrays(shapes, x, y, z, R.DT, bg)
    with _engine(pb, name, dtype) as eng:
        eng.rasterize(R.build_shapes(shapes, G), x, y, z, bg)
        for k, w in zip(("Ca", "Cb", "Da", "Db"), want):
            got = eng.download_coeffs(k).reshape(dims)
            ref_vals = g[f"{name}.{k}"]
            if dtype == "float32":
                w, ref_vals = w.astype(np.float32).astype(np.float64), ref_vals.astype(np.float32).astype(np.float64)
            assert np.array_equal(got, w), (name, dtype, k, np.abs(got - w).max())
            assert np.array_equal(got, ref_vals), (name, dtype, k)
        with pytest.raises(RuntimeError):
            eng.download_coeffs("Cby")                   # isotropic scene: no per-component arrays


@pytest.mark.gpu
def test_device_raster_each_shape_alone_equals_reference_masks():
    """Mask-level check: painting ONE entry with eps 2 over vacuum reproduces the reference's mask of that entry."""
    import prismo_b200 as pb
    from prismo_b200 import geometry as G

    g = np.load(GOLD)
    for name in ("scene3d", "scene2d"):
        dims, sp, org, shapes, bg = R.SCENES[name]
        x, y, z = R.coords(dims, sp, org)
        masks = _gold_masks(g, name, dims)
        cb_vac = kernels.coefficients(np.ones(1), np.ones(1), np.zeros(1), np.zeros(1), R.DT)[1][0]
        with _engine(pb, name, "float64") as eng:
            for n, s in enumerate(shapes):
                one = dict(s, eps_r=2.0, mu_r=1.0, sigma_e=0.0, sigma_m=0.0)
                eng.rasterize(R.build_shapes([one], G), x, y, z, (1.0, 1.0, 0.0, 0.0))
                got = eng.download_coeffs("Cb").reshape(dims) != cb_vac
                assert np.array_equal(got, masks[n]), (name, n, s["kind"], int((got != masks[n]).sum()))


@pytest.mark.gpu
def test_long_shape_list_and_empty_list():
    """> 64 entries leave the shared-memory staging path; an empty list paints the background."""
    import prismo_b200 as pb
    from prismo_b200 import geometry as G

    dims, sp, org, _, _ = R.SCENES["scene3d"]
    x, y, z = R.coords(dims, sp, org)
    rng = np.random.default_rng(5)
    shapes = [dict(kind="sphere", center=tuple(rng.uniform(-0.4, 0.4, 3)), radius=float(rng.uniform(0.05, 0.15)),
                   eps_r=float(1.5 + n * 0.1)) for n in range(70)]
    with _engine(pb, "scene3d", "float64") as eng:
        eng.rasterize(R.build_shapes(shapes, G), x, y, z, None)
        want = raster.coefficient_arrays(shapes, x, y, z, R.DT)
        assert np.array_equal(eng.download_coeffs("Cb"), want[1])
        eng.rasterize([], x, y, z, (3.0, 1.0, 0.5, 0.0))
        want = raster.coefficient_arrays([], x, y, z, R.DT, (3.0, 1.0, 0.5, 0.0))
        for k, w in zip(("Ca", "Cb", "Da", "Db"), want):
            assert np.array_equal(eng.download_coeffs(k), w)
        with pytest.raises(ValueError):
            eng.rasterize([], x[:-2], y, z)


def _seed(eng, dims, seed=11):
    import prismo_b200 as pb

    rng = np.random.default_rng(seed)
    F = {}
    for c in pb.grid.COMPONENTS:
        F[c] = rng.standard_normal(eng.field_shape(c)) * (1.0 if c[0] == "E" else 1 / 377.0)
        eng.upload(c, F[c])
    return F


@pytest.mark.gpu
def test_run_on_device_painted_coefficients_equals_host_painted_and_oracle():
    import prismo_b200 as pb
    from prismo_b200 import geometry as G

    dims, sp, org, shapes, bg = R.SCENES["scene3d"]
    x, y, z = R.coords(dims, sp, org)
    coef = raster.coefficient_arrays(shapes, x, y, z, R.DT, bg)
    with _engine(pb, "scene3d", "float64") as a, _engine(pb, "scene3d", "float64") as b:
        a.rasterize(R.build_shapes(shapes, G), x, y, z, bg)
        b.set_coeffs(*coef)
        F = _seed(a, dims)
        _seed(b, dims)
        a.run(5)
        b.run(5)
        for _ in range(5):
            kernels.step(F, coef, sp, False)
        for c in pb.grid.COMPONENTS:
            got = a.download(c)
            assert np.array_equal(got, b.download(c)) and np.array_equal(got, F[c]), c


@pytest.mark.gpu
@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_per_component_cb_sweep_equals_oracle(two_pass, dtype):
    """OPT-IN extension (parity unpinned: the reference's solver has one Cb per cell): the fused heterogeneous sweep and
    the two-pass kernels with Cb_x / Cb_y / Cb_z, from rasterised anisotropic shapes and from host arrays."""
    import prismo_b200 as pb
    from prismo_b200 import _lib
    from prismo_b200 import geometry as G

    dims, sp, org, shapes, bg = R.SCENES["aniso"]
    x, y, z = R.coords(dims, sp, org)
    Ca, Cb3, Da, Db = raster.coefficient_arrays(shapes, x, y, z, R.DT, bg)
    assert not np.array_equal(Cb3[0], Cb3[1]) and not np.array_equal(Cb3[1], Cb3[2])
    flags = _lib.FLAG_TWO_PASS if two_pass else 0
    steps = 6
    with _engine(pb, "aniso", dtype, flags=flags) as a, _engine(pb, "aniso", dtype, flags=flags) as b:
        a.rasterize(R.build_shapes(shapes, G), x, y, z, bg)
        for k, w in zip(("Ca", "Cb", "Cby", "Cbz", "Da", "Db"), (Ca, *Cb3, Da, Db)):
            w = w if dtype == "float64" else w.astype(np.float32).astype(np.float64)
            assert np.array_equal(a.download_coeffs(k), w), k
        b.set_coeffs_aniso(Ca, *Cb3, Da, Db)
        F = _seed(a, dims)
        _seed(b, dims)
        a.run(steps)
        b.run(steps)
        for _ in range(steps):
            kernels.step(F, (Ca, Cb3, Da, Db), sp, False)
        for c in pb.grid.COMPONENTS:
            got = a.download(c)
            assert np.array_equal(got, b.download(c)), c
            if dtype == "float64":
                assert np.array_equal(got, F[c]), (c, np.abs(got - F[c]).max())
            else:
                err = np.linalg.norm(got - F[c]) / np.linalg.norm(F[c])
                assert err <= 1e-4, (c, err)


@pytest.mark.gpu
def test_equal_components_reproduce_the_isotropic_sweep_and_unsupported_modes_raise():
    import prismo_b200 as pb
    from prismo_b200 import _lib

    dims, sp, org, shapes, bg = R.SCENES["scene3d"]
    x, y, z = R.coords(dims, sp, org)
    Ca, Cb, Da, Db = raster.coefficient_arrays(shapes, x, y, z, R.DT, bg)
    with _engine(pb, "scene3d", "float64") as a, _engine(pb, "scene3d", "float64") as b:
        a.set_coeffs_aniso(Ca, Cb, Cb, Cb, Da, Db)
        b.set_coeffs(Ca, Cb, Da, Db)
        _seed(a, dims)
        _seed(b, dims)
        a.run(4)
        b.run(4)
        for c in pb.grid.COMPONENTS:
            assert np.array_equal(a.download(c), b.download(c)), c
        b.set_coeffs(Ca, Cb, Da, Db)                     # back to isotropic: the extra arrays are dropped
        with pytest.raises(RuntimeError):
            b.download_coeffs("Cbz")
    with _engine(pb, "scene2d", "float64") as e2:
        d2 = R.SCENES["scene2d"][0]
        one = np.ones(d2)
        with pytest.raises(ValueError):
            e2.set_coeffs_aniso(one, one, one, one, one, one)
    with _engine(pb, "scene3d", "float64", flags=_lib.FLAG_YEE) as ey:
        with pytest.raises(ValueError):
            ey.set_coeffs_aniso(Ca, Cb, Cb, Cb, Da, Db)


@pytest.mark.gpu
def test_session_set_geometry_drives_a_simulation():
    """Session.set_geometry: the mirror Simulation runs on device-painted coefficients; equal to the same Simulation
    with host-painted material arrays."""
    import prismo_b200 as pb
    from prismo_b200 import geometry as G
    from tests import scenarios as S

    spec = dict(S.SCENARIOS["src3d_point"])
    shapes = [dict(kind="box", center=(0.5e-6, 0.5e-6, 0.4e-6), size=(0.4e-6, 2e-6, 0.3e-6), eps_r=11.9),
              dict(kind="sphere", center=(0.3e-6, 0.3e-6, 0.3e-6), radius=0.15e-6, eps_r=2.1, sigma_e=50.0)]
    sim = S.build_mirror(spec, pb, dtype="float64")
    sess = sim.solver.updater.session()
    x, y, z = G.cell_coordinates(sim.grid)
    sess.set_geometry(R.build_shapes(shapes, G), background=(1.0, 1.0, 0.0, 0.0))
    sim.run_steps(spec["steps"])
    got = S.results_mirror(sim)
    eps, mu, se, sm = raster.material_arrays(shapes, x, y, z)
    assert 0 < (eps != 1.0).sum() < eps.size
    sim2 = S.build_mirror(spec, pb, dtype="float64")
    sim2.set_materials(dict(eps_rel=eps, mu_rel=mu, sigma_e=se, sigma_m=sm))
    sim2.run_steps(spec["steps"])
    want = S.results_mirror(sim2)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["scene3d", "aniso"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("ade", [False, True])
def test_index_coded_sweep_equals_array_sweep(name, dtype, ade):
    """Material-index coding (option "het_indexed"): the sweep reads one byte per cell and looks Ca..Db up in the material
    table fdtd_rasterize wrote — the same numbers the coefficient arrays hold, so fields (and recursion state, with a
    dispersive box applied in-sweep) are bit-identical to the array path in both dtypes; fp64 also equals the oracle."""
    import prismo_b200 as pb
    from prismo_b200 import geometry as G
    from prismo_b200.engine import AdeOp

    dims, sp, org, shapes, bg = R.SCENES[name]
    x, y, z = R.coords(dims, sp, org)
    coef = raster.coefficient_arrays(shapes, x, y, z, R.DT, bg)
    steps = 7
    with _engine(pb, name, dtype) as a, _engine(pb, name, dtype) as b:
        a.set_option("het_indexed", 0)
        b.set_option("het_indexed", 1)
        ids = []
        for e in (a, b):
            e.rasterize(R.build_shapes(shapes, G), x, y, z, bg)
            if ade:
                ids.append(e.add_ade_op(AdeOp("Ez", 1, (2, 3, 1), (dims[0] - 3, dims[1] - 2, dims[2] - 4), 0.3, 0.9)))
        F = _seed(a, dims)
        _seed(b, dims)
        a.run(3); a.run(steps - 3)
        b.run(3); b.run(steps - 3)
        for c in pb.grid.COMPONENTS:
            got = b.download(c)
            assert np.array_equal(got, a.download(c)), (c, np.abs(got - a.download(c)).max())
        if ade:
            sa, sb = a.ade_state(ids[0], 0), b.ade_state(ids[1], 0)
            assert np.abs(sa).max() > 0 and np.array_equal(sa, sb)
        if dtype == "float64":
            for _ in range(steps):
                kernels.step(F, coef, sp, False)
            for c in pb.grid.COMPONENTS:
                assert np.array_equal(b.download(c), F[c]), c
        # host arrays switch the coding off again (nothing describes them as indices)
        Ca, Cb, Da, Db = coef
        if isinstance(Cb, tuple):
            b.set_coeffs_aniso(Ca, *Cb, Da, Db)
        else:
            b.set_coeffs(Ca, Cb, Da, Db)
        b.run(1)
        a.run(1)
        for c in pb.grid.COMPONENTS:
            assert np.array_equal(b.download(c), a.download(c)), c


@pytest.mark.reference
def test_oracle_raster_random_scenes_against_live_reference(ref):
    """Property-style pinning: random shape lists (all classes, all group operations, degenerate sizes, vertices and faces
    on cell coordinates) rasterised by the oracle and by the reference's own classes give identical masks, 3-D and 2-D."""
    from prismo.geometry import shapes as RS

    rng = np.random.default_rng(17)
    x = np.round(np.linspace(-0.5, 0.5, 12), 3)
    y = np.round(np.linspace(-0.4, 0.45, 11), 3)
    z3 = np.round(np.linspace(-0.3, 0.35, 9), 3)

    def rnd_prim():
        kind = rng.choice(["box", "sphere", "cylinder", "polygon"])
        c = tuple(float(v) for v in rng.choice(x, 1)) + tuple(float(v) for v in rng.choice(y, 1)) + \
            tuple(float(v) for v in rng.choice(z3, 1))                       # centres ON cell coordinates
        if kind == "box":
            size = tuple(float(v) for v in rng.choice([0.0, 0.1, 0.2, 0.3, 0.55], 3))
            return dict(kind="box", center=c, size=size)
        if kind == "sphere":
            return dict(kind="sphere", center=c, radius=float(rng.choice([0.0, 0.1, 0.25, 0.3])))
        if kind == "cylinder":
            return dict(kind="cylinder", center=c, radius=float(rng.choice([0.0, 0.1, 0.2])),
                        height=float(rng.choice([0.0, 0.2, 0.5])), axis=str(rng.choice(["x", "y", "z"])))
        n = int(rng.integers(3, 7))
        v = np.stack([rng.choice(x, n), rng.choice(y, n)], axis=1)              # vertices on cell coordinates, any winding
        zlo, zhi = sorted(float(q) for q in rng.choice(z3, 2))
        return dict(kind="polygon", vertices=[tuple(map(float, p)) for p in v], z_min=zlo, z_max=zhi)

    checked = 0
    for trial in range(40):
        z = z3 if trial % 3 else None
        if rng.random() < 0.3:
            s = dict(kind="group", operation=str(rng.choice(["union", "intersection", "difference"])),
                     shapes=[rnd_prim() for _ in range(int(rng.integers(2, 4)))])
        else:
            s = rnd_prim()
        obj = R.build_shapes([dict(s, eps_r=2.0)], RS)[0]
        want = obj.rasterize(x, y, z)
        want = want[0] if isinstance(want, tuple) else want
        got = raster.contains(s, x, y, z)
        assert np.array_equal(got.reshape(want.shape), want), (trial, s)
        checked += int(want.any())
    assert checked >= 15
