"""The oracle is pinned two ways: against the REAL reference (where /root/reference exists) and against
golden outputs of the real reference committed under tests/golden/ (everywhere, incl. the GPU box)."""
import os

import numpy as np
import pytest

from tests import scenarios as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _exact(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_oracle_matches_golden(name):
    """fp64 bit-exact for integer-free elementwise arithmetic: same operation order as the reference."""
    spec = S.SCENARIOS[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    o = S.build_oracle(spec)
    o.run_steps(spec["steps"])
    res = S.results_oracle(o)
    assert sorted(res) == sorted(gold.files)
    for k in gold.files:
        assert _exact(res[k], gold[k]), f"{name}:{k} rel-L2 {S.rel_l2(res[k], gold[k]):.3e}"


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_oracle_matches_live_reference(name, ref):
    spec = S.SCENARIOS[name]
    r = S.build_reference(spec, ref)
    S.step_reference(r, spec["steps"])
    o = S.build_oracle(spec)
    o.run_steps(spec["steps"])
    rr, oo = S.results_reference(r), S.results_oracle(o)
    assert sorted(rr) == sorted(oo)
    for k in rr:
        assert _exact(oo[k], rr[k]), f"{name}:{k} rel-L2 {S.rel_l2(oo[k], rr[k]):.3e}"


@pytest.mark.reference
@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_golden_is_current(name, ref):
    """The committed fixtures are what the reference produces today (guards against stale goldens)."""
    spec = S.SCENARIOS[name]
    r = S.build_reference(spec, ref)
    S.step_reference(r, spec["steps"])
    rr = S.results_reference(r)
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    assert sorted(rr) == sorted(gold.files)
    for k in rr:
        assert _exact(rr[k], gold[k]), f"{name}:{k}"


def test_scenarios_are_not_trivial():
    """Sources inject, gates flip and monitors see non-zero data in the fixtures."""
    g = np.load(os.path.join(GOLD, "src3d_point.npz"))
    assert np.abs(g["F_Ez"]).max() > 0 and np.abs(g["F_Hy"]).max() > 0
    # the first-step gate matters: with Ex = Ey = 0 and sigma_m != 0 the reference skips the Hz decay once
    from oracle import kernels
    spec = S.SCENARIOS["upd2d_gate_tm"]
    o = S.build_oracle(spec)
    hz0 = o.F["Hz"].copy()
    kernels.update_h(o.F, o.coeffs[2], o.coeffs[3], o.grid.spacing, True)
    assert np.array_equal(o.F["Hz"], hz0) and np.abs(hz0).max() > 0 and o.coeffs[2].max() < 1.0
    g = np.load(os.path.join(GOLD, "mon2d_all.npz"))
    assert np.abs(g["m1_dft_Ey"]).max() > 0 and np.abs(g["m4_ct_0"]).max() > 0 and len(g["m2_p"]) == 8


def test_reference_vacuum_coefficients():
    """The exact values the reference's own tests pin (tests/test_core_solver.py:140-176)."""
    from oracle import kernels

    dt = 1.3e-17
    one = np.ones((3, 3, 3))
    Ca, Cb, Da, Db = kernels.coefficients(one, one, 0 * one, 0 * one, dt)
    assert np.all(Ca == 1.0) and np.all(Da == 1.0)
    assert np.all(Cb == dt / 8.854187817e-12) and np.all(Db == dt / (4 * np.pi * 1e-7))
    Ca, Cb, _, _ = kernels.coefficients(11.9 * one, one, 0 * one, 0 * one, dt)
    assert np.allclose(Cb, dt / (8.854187817e-12 * 11.9), rtol=1e-15)
