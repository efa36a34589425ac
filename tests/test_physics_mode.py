"""Opt-in physics mode (stable Yee leap-frog + CPML).  PARITY UNPINNED — the reference's scheme is unstable and
its CPML is never applied — so this is checked against our own CPU restatement (oracle/yee.py) and by physics:
bounded energy, and > 40 dB less reflected energy than the same run without the layer."""
import numpy as np
import pytest

import prismo_b200 as pb
from prismo_b200 import _lib, cpml

C0 = 299792458.0
EPS0, MU0 = 8.854187817e-12, 4 * np.pi * 1e-7
COMPS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


def test_cpml_profiles_are_identity_outside_the_layers():
    p = cpml.PMLParams(thickness=8)
    c = cpml.axis_coefficients(64, 2e-8, 3e-17, p)
    assert c.shape == (6, 64)
    mid = slice(9, 54)
    assert np.all(c[0, mid] == 0) and np.all(c[1, mid] == 0) and np.all(c[2, mid] == 1)
    assert np.all(c[3, mid] == 0) and np.all(c[4, mid] == 0) and np.all(c[5, mid] == 1)
    assert 0 < c[0, 0] < 1 and c[1, 0] < 0 and 0 < c[2, 0] < 1          # b in (0,1), a < 0, 1/kappa < 1 in the layer
    assert np.allclose(c[3, :8], c[3, -1:-9:-1])                          # symmetric layers
    tab = cpml.coefficient_table((20, 24, 28), (2e-8,) * 3, 3e-17, cpml.PMLParams(thickness=4))
    assert tab.size == 6 * (20 + 24 + 28)


def _energy(f, d):
    e = sum(float(np.sum(f[c].astype(np.float64) ** 2)) for c in COMPS[:3])
    h = sum(float(np.sum(f[c].astype(np.float64) ** 2)) for c in COMPS[3:])
    return 0.5 * (EPS0 * e + MU0 * h) * d ** 3


@pytest.mark.gpu
@pytest.mark.parametrize("thickness", [0, 6])
def test_physics_mode_matches_its_oracle(thickness):
    from oracle.yee import YeeOracle

    dims, d = (28, 24, 26), 2e-8
    dt = 0.9 * d / (C0 * np.sqrt(3))
    params = cpml.PMLParams(thickness=thickness, alpha_max=0.05)
    eng = pb.Engine(3, dims, (d,) * 3, dt, dtype="float64", flags=_lib.FLAG_YEE)
    axes = None
    if thickness:
        eng.set_cpml(thickness, cpml.coefficient_table(dims, (d,) * 3, dt, params))
        axes = [cpml.axis_coefficients(n, d, dt, params) for n in dims]
    o = YeeOracle(dims, (d,) * 3, dt, (1.0, dt / EPS0, 1.0, dt / MU0), axes)
    rng = np.random.default_rng(1)
    for c in COMPS:
        o.F[c][...] = rng.standard_normal(o.F[c].shape) * (1.0 if c[0] == "E" else 1 / 377.0)
        eng.upload(c, o.F[c])
    eng.run(40)
    for _ in range(40):
        o.step()
    for c in COMPS:
        got = eng.download(c)
        assert np.array_equal(got, o.F[c]), f"{c}: max rel {np.abs(got - o.F[c]).max() / np.abs(o.F[c]).max():.2e}"
    eng.close()


@pytest.mark.gpu
def test_physics_mode_is_stable_and_cpml_absorbs():
    """A compact pulse in a 72^3 box.  (1) The leap-frog is stable: without a layer the energy stays within a
    few % of its initial value over 600 steps (the reference's scheme gains 34 orders of magnitude in 50 steps).
    (2) With a 10-cell CPML less than 1e-4 of the energy (-40 dB) is left once the pulse has crossed the box."""
    n, d = 72, 2e-8
    dt = 0.9 * d / (C0 * np.sqrt(3))
    x = (np.arange(n) - n / 2)[:, None, None]
    y = (np.arange(n) - n / 2)[None, :, None]
    z = (np.arange(n) - n / 2)[None, None, :]
    pulse = np.exp(-(x ** 2 + y ** 2 + z ** 2) / (2 * 4.0 ** 2))
    left = {}
    for t in (0, 10):
        eng = pb.Engine(3, (n, n, n), (d,) * 3, dt, dtype="float32", flags=_lib.FLAG_YEE)
        if t:
            eng.set_cpml(t, cpml.coefficient_table((n, n, n), (d,) * 3, dt, cpml.PMLParams(thickness=t)))
        eng.upload("Ez", pulse[:-1, :-1, :])
        e0 = _energy({c: eng.download(c) for c in COMPS}, d)
        eng.run(600)
        f = {c: eng.download(c) for c in COMPS}
        assert all(np.isfinite(a).all() for a in f.values())
        left[t] = _energy(f, d) / e0
        eng.close()
    assert 0.5 < left[0] < 1.5, left          # closed box: energy conserved (leap-frog energy oscillates slightly)
    assert left[10] < 1e-4, left              # open box: > 40 dB absorbed
