"""The plugin path on a real GPU: ``prismo.set_backend("b200")`` + the UNMODIFIED reference ``Simulation`` / sources /
monitors (staged copy under oracle/_ref, see oracle/stage_reference.py) driving the real libfdtd_b200.so — what
BASELINE.json's north_star describes (backends/backend_manager.py:139-186, core/simulation.py:107-164).
fp64 (the plugin's default dtype) must reproduce the reference goldens bit for bit."""
import os

import numpy as np
import pytest

import prismo_b200 as pb
from tests import scenarios as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"src3d_mode": 1e-14}


def _compare(name, res, gold):
    assert sorted(res) == sorted(gold)
    for k in gold:
        if name in TOL:
            assert S.rel_l2(res[k], gold[k]) <= TOL[name], f"{name}:{k}"
        else:
            assert res[k].shape == gold[k].shape and np.array_equal(res[k], gold[k], equal_nan=True), \
                f"{name}:{k} rel-L2 {S.rel_l2(res[k], gold[k]):.3e}"


@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_plugin_runs_reference_objects_on_the_gpu(name, ref):
    pb.register()
    spec = S.SCENARIOS[name]
    gold = dict(np.load(os.path.join(GOLD, name + ".npz")))
    try:
        sim = S.build_reference(spec, ref, backend="b200")
        assert sim.solver.updater.backend.name == "b200" and sim.fields.backend.is_gpu
        sim.step()                                              # Simulation.step
        sim.run((spec["steps"] - 1 - 0.5) * sim.dt)             # Simulation.run: ceil -> steps-1 more
        sess = sim.solver.updater._b200_session
        assert type(sess.engine).__module__.startswith("prismo_b200") and sess.engine.kernel_launches > 0
        _compare(name, S.results_reference(sim), gold)
    finally:
        ref.set_backend("numpy")


def test_plugin_run_with_progress_callback_and_solver_entry_points(ref):
    """Simulation.run(progress_callback=...) chunks the device run at the reference's callback points; FDTDSolver.step /
    run_steps and MaxwellUpdater.step go through the engine too; all equal the stock NumPy backend bit for bit."""
    pb.register()
    spec = S.SCENARIOS["src3d_tfsf"]
    try:
        want = S.build_reference(spec, ref, backend="numpy")
        calls_w = []
        want.run((spec["steps"] - 0.5) * want.dt, progress_callback=lambda *a: calls_w.append(a[:2]), progress_interval=3)
        got = S.build_reference(spec, ref, backend="b200")
        calls_g = []
        got.run((spec["steps"] - 0.5) * got.dt, progress_callback=lambda *a: calls_g.append(a[:2]), progress_interval=3)
        assert calls_g == calls_w and got.step_count == want.step_count == spec["steps"]
        assert got.current_time == want.current_time
        for c in S.COMPONENTS:
            assert np.array_equal(got.fields[c], want.fields[c]), c
        # solver-level entry points
        a = S.build_reference(S.SCENARIOS["upd3d_het"], ref, backend="numpy")
        b = S.build_reference(S.SCENARIOS["upd3d_het"], ref, backend="b200")
        for s in (a, b):
            s.solver.step(s.fields)
            s.solver.run_steps(3)
            s.solver.updater.step(s.fields)
            s.solver.updater.update_magnetic_fields(s.fields)
            s.solver.updater.update_electric_fields(s.fields)
        for c in S.COMPONENTS:
            assert np.array_equal(a.fields[c], b.fields[c]), c
        assert a.solver.step_count == b.solver.step_count and a.solver.time == b.solver.time
    finally:
        ref.set_backend("numpy")


def test_backend_object_is_wired_to_the_library(ref):
    pb.register()
    try:
        b = ref.set_backend("b200")
        info = b.get_memory_info()
        assert info["backend"] == "b200" and info["total_bytes"] > 100e9 > 0 and 0 < info["free_bytes"] <= info["total_bytes"]
        b.synchronize()
        z = b.zeros((4, 5))
        assert isinstance(z, np.ndarray) and z.dtype == np.float64 and not z.any()
    finally:
        ref.set_backend("numpy")


def test_plugin_set_geometry_equals_the_reference_rasterise_pipeline(ref):
    """Row f4: the reference's OWN shape objects painted on the device (pb.set_geometry) against the reference's host
    pipeline — Shape.rasterize on grid coordinates, eps_rel[mask] = epsilon_r in list order, FDTDSolver(material_arrays)
    on the stock NumPy backend.  Fields after the run are bit-identical."""
    from prismo.core.solver import FDTDSolver
    from prismo.geometry import shapes as RS

    pb.register()
    spec = S.SCENARIOS["src3d_point"]
    shapes = [RS.Box(RS.Material("Si", 11.9), (0.3e-6, 0.3e-6, 0.15e-6), (0.4e-6, 2e-6, 0.2e-6)),
              RS.Sphere(RS.Material("glass", 2.1), (0.2e-6, 0.35e-6, 0.4e-6), 0.17e-6),
              RS.Cylinder(RS.Material("rod", 4.0, 1.5), (0.4e-6, 0.2e-6, 0.3e-6), 0.1e-6, 0.45e-6, "y"),
              RS.GeometryGroup([RS.Box(RS.Material("ring", 6.0), (0.1e-6, 0.1e-6, 0.1e-6), (0.3e-6, 0.3e-6, 0.3e-6)),
                                RS.Sphere(RS.Material("hole", 1.0), (0.1e-6, 0.1e-6, 0.1e-6), 0.1e-6)], "difference")]
    try:
        want = S.build_reference(spec, ref, backend="numpy")
        g = want.grid
        x, y, z = (g.origin[d] + np.arange(g.dimensions[d]) * g.spacing[d] for d in range(3))
        eps, mu = np.ones(g.dimensions), np.ones(g.dimensions)
        for sh in shapes:
            m = sh.rasterize(x, y, z)
            mat = sh.shapes[0].material if isinstance(m, tuple) else sh.material
            m = m[0] if isinstance(m, tuple) else m
            assert 0 < m.sum() < m.size
            eps[m], mu[m] = mat.epsilon_r, mat.mu_r
        want.solver = FDTDSolver(want.grid, want.dt, dict(eps_rel=eps, mu_rel=mu, sigma_e=0 * eps, sigma_m=0 * eps))
        S.step_reference(want, spec["steps"])
        got = S.build_reference(spec, ref, backend="b200")
        pb.set_geometry(got, shapes)
        got.step()
        got.run((spec["steps"] - 1 - 0.5) * got.dt)
        for c in S.COMPONENTS:
            assert np.array_equal(got.fields[c], want.fields[c]), c
        sess = got.solver.updater._b200_session
        assert np.array_equal(sess.engine.download_coeffs("Cb"), want.solver.updater.Cb)
        pb.clear_geometry(got)                                   # back to the updater's own (vacuum) arrays
        got.step()
        with pytest.raises(RuntimeError):                        # vacuum again: uniform coefficients, no arrays on the device
            sess.engine.download_coeffs("Cb")
    finally:
        ref.set_backend("numpy")
