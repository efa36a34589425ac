"""The ALGORITHM of the opt-in fused physics sweep (csrc/fdtd_yee_fused.cuh), modelled statement by statement in NumPy
(tools/yee_fused_model.py), must equal the physics-mode oracle (oracle/yee.py) bit for bit: tiling with rims, x-segment
prologues, first-row / first-lane global fetches, update ranges, CPML slab offsets and the psi ping-pong.  The CUDA
kernel itself is checked on a GPU by tools/check_yee_fused.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

from oracle.yee import YeeOracle  # noqa: E402
from prismo_b200 import cpml  # noqa: E402
from yee_fused_model import Model  # noqa: E402

C0 = 299792458.0
EPS0, MU0 = 8.854187817e-12, 4 * np.pi * 1e-7
COMPS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


@pytest.mark.parametrize("dims,thickness,tiling", [
    ((9, 8, 11), 0, dict(TJ=3, W=4, own_lanes=2, V=2, lx=4)),
    ((9, 8, 11), 2, dict(TJ=3, W=4, own_lanes=2, V=2, lx=4)),
    ((12, 10, 9), 3, dict(TJ=4, W=8, own_lanes=6, V=2, lx=5)),
    ((7, 13, 10), 2, dict(TJ=15, W=8, own_lanes=6, V=4, lx=32)),      # one segment, one row tile, rim lanes only
    ((10, 7, 17), 3, dict(TJ=2, W=4, own_lanes=2, V=1, lx=3)),
])
def test_fused_yee_algorithm_equals_the_oracle(dims, thickness, tiling):
    d = (2e-8, 2.5e-8, 3e-8)
    dt = 0.9 / (C0 * np.sqrt(sum(1 / s ** 2 for s in d)))
    coeffs = (1.0, dt / EPS0, 1.0, dt / MU0)
    axes = None
    if thickness:
        params = cpml.PMLParams(thickness=thickness, alpha_max=0.05)
        axes = [cpml.axis_coefficients(n, s, dt, params) for n, s in zip(dims, d)]
    o = YeeOracle(dims, d, dt, coeffs, axes)
    m = Model(dims, d, dt, coeffs, axes, thickness, **tiling)
    rng = np.random.default_rng(3)
    for c in COMPS:
        o.F[c][...] = rng.standard_normal(o.F[c].shape) * (1.0 if c[0] == "E" else 1 / 377.0)
        m.upload(c, o.F[c])
    for step in range(5):
        o.step()
        m.step()
        for c in COMPS:
            got = m.download(c, o.F[c].shape)
            assert np.array_equal(got, o.F[c]), f"step {step} {c}: max |diff| {np.abs(got - o.F[c]).max():.3e}"
    # padding stays zero (cells outside the staggered shapes are copied through untouched)
    for c in COMPS:
        A = m.F[m.cur][c.lower()]
        s = o.F[c].shape
        assert not A[s[0]:].any() and not A[:, s[1]:].any() and not A[:, :, s[2]:].any()
