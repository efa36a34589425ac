"""Host-side logic without a GPU: lowering, tables, chunking, write-back and the prismo plugin, run on a
NumPy test double of the engine (tests/_fake_engine.py) and compared with golden / live reference results."""
import os

import numpy as np
import pytest

import prismo_b200 as pb
import prismo_b200.session as session
from tests import scenarios as S
from tests._fake_engine import FakeEngine

GOLD = os.path.join(os.path.dirname(__file__), "golden")
# scipy.ndimage.zoom(a*P) vs a*zoom(P): the one place the lowering is not bitwise (a few ulp)
TOL = {"src3d_mode": 1e-14}


@pytest.fixture()
def fake(monkeypatch):
    monkeypatch.setattr(session, "Engine", FakeEngine)
    FakeEngine.instances.clear()
    return FakeEngine


def _compare(name, res, gold):
    assert sorted(res) == sorted(gold)
    for k in gold:
        if name in TOL:
            assert S.rel_l2(res[k], gold[k]) <= TOL[name], f"{name}:{k}"
        else:
            assert res[k].shape == gold[k].shape and np.array_equal(res[k], gold[k], equal_nan=True), \
                f"{name}:{k} rel-L2 {S.rel_l2(res[k], gold[k]):.3e}"


@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_mirror_classes_lower_like_reference(name, fake):
    spec = S.SCENARIOS[name]
    gold = dict(np.load(os.path.join(GOLD, name + ".npz")))
    sim = S.build_mirror(spec, pb)
    sim.run_steps(3)                              # two submissions: DFT sums and records must carry over
    sim.run_steps(spec["steps"] - 3)
    _compare(name, S.results_mirror(sim), gold)
    assert sim.step_count == spec["steps"] and sim.solver.step_count == spec["steps"]


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_plugin_runs_reference_objects(name, fake, ref):
    """prismo.set_backend('b200') + unchanged reference Simulation/sources/monitors."""
    pb.register()
    spec = S.SCENARIOS[name]
    gold = dict(np.load(os.path.join(GOLD, name + ".npz")))
    try:
        sim = S.build_reference(spec, ref, backend="b200")
        assert sim.solver.updater.backend.name == "b200" and sim.fields.backend.is_gpu
        sim.step()
        sim.run((spec["steps"] - 1 - 0.5) * sim.dt)        # ceil -> steps-1 more (attached ADE runs inside)
        assert len(fake.instances) == 1                    # one engine per simulation, reused
        _compare(name, S.results_reference(sim), gold)
    finally:
        ref.set_backend("numpy")


@pytest.mark.reference
def test_plugin_leaves_other_backends_alone(fake, ref):
    pb.register()
    assert "b200" in ref.list_available_backends() and "numpy" in ref.list_available_backends()
    with pytest.raises(ValueError, match="Unknown backend"):
        ref.set_backend("nope")
    ref.set_backend("numpy")
    sim = S.build_reference(S.SCENARIOS["upd3d_vac"], ref)
    sim.step()
    assert not fake.instances                              # stock NumPy path, no engine created
    assert ref.get_backend().name == "numpy"
    b = ref.get_backend("b200")
    assert b.name == "b200" and b.is_gpu and isinstance(b.zeros((2, 2)), np.ndarray)
    ref.set_backend("numpy")


@pytest.mark.reference
def test_plugin_progress_callback_and_solver_entry_points(fake, ref):
    pb.register()
    try:
        spec = S.SCENARIOS["src3d_plane"]
        ref.set_backend("numpy")
        a = S.build_reference(spec, ref)
        calls_a = []
        a.run(25.5 * a.dt, progress_callback=lambda i, n, t, el: calls_a.append((i, n, t)), progress_interval=7)
        b = S.build_reference(spec, ref, backend="b200")
        calls_b = []
        b.run(25.5 * b.dt, progress_callback=lambda i, n, t, el: calls_b.append((i, n, t)), progress_interval=7)
        assert calls_a == calls_b and len(calls_a) == 5
        for c in S.COMPONENTS:
            assert np.array_equal(a.fields[c], b.fields[c], equal_nan=True)
        # FDTDSolver.run_steps with a per-step callback, MaxwellUpdater half steps
        seen = []
        b.solver.run_steps(3, callback=lambda s, k: seen.append((k, s.step_count)))
        a.solver.fields = a.fields
        assert [k for k, _ in seen] == [0, 1, 2]
        b.solver.updater.update_magnetic_fields(b.fields)
        b.solver.updater.update_electric_fields(b.fields)
    finally:
        ref.set_backend("numpy")


def test_tables_follow_accumulated_time(fake):
    t = session.Session.step_times(0.0, 0.1, 5)
    acc = 0.0
    for k in range(5):
        acc += 0.1
        assert t[k] == acc                               # repeated adds, not (k+1)*dt
    assert session.Session.step_times(0.0, 0.1, 10)[-1] != 10 * 0.1     # 0.9999999999999999


def test_record_pool_chunking(fake, monkeypatch):
    """A tiny record pool forces many chunks; results must not change."""
    spec = S.SCENARIOS["mon3d_field"]
    gold = dict(np.load(os.path.join(GOLD, "mon3d_field.npz")))
    monkeypatch.setattr(session, "_RECORD_POOL_BYTES", 3 * 8 * 6 * 200)
    sim = S.build_mirror(spec, pb)
    sim.run_steps(spec["steps"])
    _compare("mon3d_field", S.results_mirror(sim), gold)


def test_unknown_source_is_rejected(fake):
    class Weird:
        enabled = True

    sim = pb.Simulation((0.5e-6, 0.5e-6, 0.5e-6), 20e6, pml_layers=2)
    sim.sources.append(Weird())
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        sim.step()


def test_reference_error_behaviour(fake):
    g = pb.YeeGrid(pb.GridSpec((0.5e-6, 0.5e-6, 0.5e-6), 20e6, 2))
    with pytest.raises(ValueError, match="Courant"):
        pb.MaxwellUpdater(g, dt=1.0)                               # solver.py:67-71
    f = pb.ElectromagneticFields(g)
    with pytest.raises(KeyError):
        f["Bx"]                                                    # fields.py:86-87
    with pytest.raises(ValueError, match="Shape mismatch"):
        f["Ex"] = np.zeros((2, 2, 2))
    with pytest.raises(ValueError):
        pb.PlaneWaveSource((0, 0, 0), (0, 0, 0), "x", "x", 1e14, pulse=False)
    with pytest.raises(ValueError, match="pulse_width"):
        pb.ElectricDipole((0, 0, 0), "x", 1e14)
    sim2 = pb.Simulation((1e-6, 0.8e-6, 0.0), 20e6, pml_layers=3)
    sim3 = pb.Simulation((0.5e-6, 0.5e-6, 0.5e-6), 20e6, pml_layers=2)
    sim3.add_monitor(pb.DFTMonitor((0, 0, 0), (0, 0, 0), [1e14]))
    with pytest.raises(ValueError, match="2-D only"):              # reference raises in 3-D too (F8)
        sim3.step()
    sim2.add_source(pb.ModeSource((0, 0, 0), (0, 0.2e-6, 0), None, "+x", None))
    with pytest.raises(IndexError):
        sim2.step()


def test_grid_matches_reference_quirks():
    g = pb.YeeGrid(pb.GridSpec((10e-6, 5e-6, 0.0), 50e6, 10))
    assert g.dimensions == (521, 271, 1) and g.is_2d            # float ceil: 10e-6*50e6 -> 501 (+20)
    assert g.get_field_shape("Ex") == (521, 270) and g.get_field_shape("Hz") == (521, 271)
    assert g.point_to_index((-1.0, 1.0, 0.0)) == (0, g.Ny - 1, 0)   # clamps to the PHYSICAL count
    g3 = pb.YeeGrid(pb.GridSpec((1e-6, 1e-6, 1e-6), 20e6, 5))
    assert g3.get_field_shape("Ey") == (29, 30, 29)
    ix = g3.get_component_indices("Ez", 0, 100, 3, 4, 0, 100)
    assert ix[0].shape == (20, 1, 1) and ix[1].ravel().tolist() == [3] and ix[2].shape == (1, 1, 20)


def test_region_correct_flux_monitor_and_s_parameter_helpers(fake):
    """Extension (off by default): FluxMonitor(region_correct=True) integrates over its real surface in 3-D and its
    six DFTs cover the same box, so P(w) from the reference's formula and mode coefficients come from real planes."""
    import types

    from prismo_b200 import postprocess as pp

    spec = dict(S.SCENARIOS["mon3d_field"], monitors=[])
    sim = S.build_mirror(spec, pb)
    fm = pb.FluxMonitor((0.3e-6, 0.25e-6, 0.2e-6), (0.0, 0.3e-6, 0.2e-6), "x", frequencies=[S.F0, 1.1 * S.F0],
                        region_correct=True)
    sim.add_monitor(fm)
    sim.run_steps(4)
    sim.run_steps(3)
    t, p = fm.get_time_domain_power()
    assert len(t) == 7 and len(p) == 7 and np.all(np.isfinite(p)) and np.abs(p).max() > 0
    assert fm._dft_ey.shape[0] == 2 and fm._dft_ey.ndim == 4 and fm._dft_ey.shape[1] == 1
    pw = fm.get_frequency_domain_power()
    assert pw.shape == (2,) and np.all(np.isfinite(pw))
    # the device-side reduction equals the same sum formed from the fields after the last step
    i = sim.grid.point_to_index((0.3e-6, 0.25e-6, 0.2e-6))[0]
    box = (slice(i, i + 1),) + tuple(slice(0, s) for s in fm._dft_ey.shape[2:])
    sl = tuple(slice(b.start, b.stop) for b in box)
    g = sim.grid
    lo = [g.component_box(c, *g.region_bounds(fm.center, fm.size)) for c in S.COMPONENTS]
    common = tuple(slice(max(b[a][0] for b in lo), min(b[a][1] for b in lo)) for a in range(3))
    ey, hz, ez, hy = (sim.fields[c][common] for c in ("Ey", "Hz", "Ez", "Hy"))
    want = float(np.sum(ey * hz - ez * hy)) * g.dy * g.dz
    assert abs(p[-1] - want) <= 1e-12 * max(abs(want), np.abs(ey * hz).sum() * g.dy * g.dz)
    mode = types.SimpleNamespace(**S.make_mode(8, 8, 5))
    a = pp.mode_coefficient_from_dft(fm, mode, 0)
    assert np.isfinite(a) and pp.s_parameter(a, a) == 1


# ---- field residency (session.ResidentFields) --------------------------------------------------------------------
def test_fields_stay_resident_between_advances(fake):
    """A step loop that does not look at the fields moves them once up and never down; reading one component brings
    that component down (into the SAME array object) and sends it up again with the next advance; results are the
    ones of the push-all / pull-all protocol."""
    spec = S.SCENARIOS["src3d_plane"]
    gold = dict(np.load(os.path.join(GOLD, "src3d_plane.npz")))
    sim = S.build_mirror(spec, pb)
    ez_obj = sim.fields._fields["Ez"] if not isinstance(sim.fields._fields, session.ResidentFields) else sim.fields._fields._raw("Ez")
    sess = sim.solver.updater.session()
    for _ in range(3):
        sim.step()
    assert (sess.h2d_arrays, sess.d2h_arrays) == (6, 0)
    assert isinstance(sim.fields._fields, session.ResidentFields) and sim.fields._fields.stale == set(S.COMPONENTS)
    ez = sim.fields["Ez"]                                   # lazy download of ONE component
    assert ez is ez_obj and (sess.h2d_arrays, sess.d2h_arrays) == (6, 1)
    assert sim.fields.Ez is ez and sess.d2h_arrays == 1     # second look: already fresh
    sim.run_steps(spec["steps"] - 3)
    assert (sess.h2d_arrays, sess.d2h_arrays) == (7, 1)     # only the array that was handed out went up again
    _compare("src3d_plane", S.results_mirror(sim), gold)
    assert sess.d2h_arrays == 7


def test_user_writes_between_advances_are_honoured(fake):
    spec = S.SCENARIOS["upd3d_vac"]
    a, b = S.build_mirror(spec, pb), S.build_mirror(spec, pb)
    for sim in (a, b):
        sim.run_steps(2)
    a.fields["Ey"][3, 2, 1] += 0.25                       # write through the handed-out array
    b.fields["Ey"] = np.array(b.fields["Ey"]) ; b.fields["Ey"][3, 2, 1] += 0.25     # __setitem__ path
    for sim in (a, b):
        sim.run_steps(2)
    for c in S.COMPONENTS:
        assert np.array_equal(a.fields[c], b.fields[c])
    ref = S.build_mirror(spec, pb)
    ref.run_steps(4)
    assert not np.array_equal(ref.fields["Hz"], a.fields["Hz"])


def test_closing_a_session_brings_the_host_up_to_date(fake):
    spec = S.SCENARIOS["upd3d_vac"]
    a, b = S.build_mirror(spec, pb), S.build_mirror(spec, pb)
    a.run_steps(3)
    b.run_steps(3)
    a.solver.updater.session().close()
    assert not a.fields._fields.stale
    for c in S.COMPONENTS:
        assert np.array_equal(a.fields._fields._raw(c), b.fields[c])


@pytest.mark.reference
def test_amplitudes_feed_the_reference_sparameter_analyzer_and_touchstone(ref, tmp_path):
    """Row f2's last hop: forward / backward amplitudes -> the reference's own SParameterAnalyzer.add_mode_data
    (analysis/sparameters.py:88-122) -> export_touchstone (:233); S11 = b1 / a1, S21 = a2 / a1 as two_port_s_parameters."""
    from prismo.analysis.sparameters import SParameterAnalyzer, export_touchstone

    from prismo_b200 import postprocess as P

    rng = np.random.default_rng(2)
    freqs = np.array([1.8e14, 1.9e14, 2.0e14])
    lam = 299792458.0 / freqs
    a1, b1, a2 = (rng.standard_normal(3) + 1j * rng.standard_normal(3) for _ in range(3))
    phi = 2 * np.pi * 2.4 / lam * 0.4e-6
    left1, right1 = a1 + b1, a1 * np.exp(1j * phi) + b1 * np.exp(-1j * phi)            # what the two planes of port 1 see
    f1, g1 = zip(*[P.separate_forward_backward(l, r, 2.4, 0.4e-6, w) for l, r, w in zip(left1, right1, lam)])
    assert np.allclose(f1, a1, rtol=1e-12) and np.allclose(g1, b1, rtol=1e-12)
    an = SParameterAnalyzer(2, freqs)
    P.fill_sparameter_analyzer(an, 0, {0: (np.array(f1), np.array(g1)), 1: (a2, np.zeros(3))})
    assert np.allclose(an.get_s_parameter(0, 0), b1 / a1, rtol=1e-12)
    assert np.allclose(an.get_s_parameter(1, 0), a2 / a1, rtol=1e-12)
    path = tmp_path / "coupler.s2p"
    export_touchstone(path, freqs, an.s_matrix)
    rows = [ln.split() for ln in open(path) if ln[0] not in "!#"]
    assert len(rows) == 3 and len(rows[0]) == 1 + 2 * 4
    s21 = np.array([float(r[5]) + 1j * float(r[6]) for r in rows])                  # row-major: S11, S12, S21, S22
    assert np.allclose(s21, a2 / a1, rtol=1e-9)


@pytest.mark.reference
def test_plugin_set_geometry_with_reference_shapes(fake, ref):
    """pb.set_geometry(sim, [reference shape objects]) against the reference's own host pipeline (Shape.rasterize on the
    grid coordinates, eps_rel[mask] = ..., FDTDSolver(material_arrays)) on the stock NumPy backend; the mirror
    Simulation.set_geometry with mirror shapes gives the same fields.  (Engine double: the rasterisation oracle paints.)"""
    from prismo.core.solver import FDTDSolver
    from prismo.geometry import shapes as RS

    from prismo_b200 import geometry as G

    pb.register()
    spec = S.SCENARIOS["src3d_point"]

    def shapes(mod):
        return [mod.Box(mod.Material("Si", 11.9), (0.3e-6, 0.3e-6, 0.15e-6), (0.4e-6, 2e-6, 0.2e-6)),
                mod.Sphere(mod.Material("glass", 2.1), (0.2e-6, 0.35e-6, 0.4e-6), 0.17e-6),
                mod.Cylinder(mod.Material("rod", 4.0, 1.5), (0.4e-6, 0.2e-6, 0.3e-6), 0.1e-6, 0.45e-6, "y")]

    try:
        want = S.build_reference(spec, ref, backend="numpy")
        g = want.grid
        x, y, z = (g.origin[d] + np.arange(g.dimensions[d]) * g.spacing[d] for d in range(3))
        eps, mu = np.ones(g.dimensions), np.ones(g.dimensions)
        for sh in shapes(RS):
            m = sh.rasterize(x, y, z)
            assert 0 < m.sum() < m.size
            eps[m], mu[m] = sh.material.epsilon_r, sh.material.mu_r
        want.solver = FDTDSolver(want.grid, want.dt, dict(eps_rel=eps, mu_rel=mu, sigma_e=0 * eps, sigma_m=0 * eps))
        S.step_reference(want, spec["steps"])
        got = S.build_reference(spec, ref, backend="b200")
        pb.set_geometry(got, shapes(RS))
        got.step()
        got.run((spec["steps"] - 1 - 0.5) * got.dt)
        for c in S.COMPONENTS:
            assert np.array_equal(got.fields[c], want.fields[c]), c
        with pytest.raises(RuntimeError):
            pb.set_geometry(want, shapes(RS))              # a NumPy-backend simulation has no device to paint on
    finally:
        ref.set_backend("numpy")
    mirror = S.build_mirror(spec, pb)
    mirror.set_geometry(shapes(G))
    mirror.run_steps(spec["steps"])
    for c in S.COMPONENTS:
        assert np.array_equal(mirror.fields[c], want.fields[c]), c
