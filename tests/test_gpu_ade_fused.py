"""Dispersive-medium recursions INSIDE the fused sweeps (north-star subsystem 4).

1. Uncoupled (the reference's behaviour, materials/ade.py:116-160, :291-308): the recursion of step n is applied by the
   sweep of step n+1 on its E-stage input registers; the polarisation state must equal what the separate k_ade kernel
   produces after every step, bit for bit, through graph replays, odd tails and a source that overlaps the medium.
2. Coupled (opt-in extension, PARITY UNPINNED: the reference never feeds P back): against oracle/ade.py::coupled_step,
   fp64 bitwise, fp32 within 1e-4.
"""
import numpy as np
import pytest

import prismo_b200 as pb
from oracle import ade as OA
from oracle import kernels
from prismo_b200.engine import AdeOp

pytestmark = pytest.mark.gpu
COMPS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
D = 2e-8
DT = 0.5 * D / (299792458.0 * np.sqrt(3))
EPS0, MU0 = 8.854187817e-12, 4 * np.pi * 1e-7
LORENTZ = [(2 * np.pi * 2.5e14, 1.0, 1e13), (2 * np.pi * 4e14, 0.4, 3e13)]


def _eps(dims, rng):
    eps = np.ones(dims)
    eps[:, dims[1] // 3: 2 * dims[1] // 3, dims[2] // 4: dims[2] // 2] = 12.11
    eps[:, :, : dims[2] // 4] = 2.07
    return eps + 0.01 * rng.random(dims)


def _make(dims, dtype, het, rng):
    eng = pb.Engine(3, dims, (D,) * 3, DT, dtype=dtype)
    if het:
        eps = _eps(dims, rng)
        one = np.ones(dims)
        eng.set_coeffs(one, DT / (EPS0 * eps), one, one * (DT / MU0))
        coeffs = (one, DT / (EPS0 * eps), one, one * (DT / MU0))
    else:
        eng.set_uniform_coeffs(1.0, DT / EPS0, 1.0, DT / MU0)
        one = np.ones(dims)
        coeffs = (one, one * (DT / EPS0), one, one * (DT / MU0))
    return eng, coeffs


def _ops(eng):
    """Lorentz (2 poles, masked) on Ez, Drude on Ex, Debye on Ey, over boxes that cross tile rims and x-segments."""
    sz, sx, sy = eng.field_shape("Ez"), eng.field_shape("Ex"), eng.field_shape("Ey")
    rng = np.random.default_rng(5)
    box_z = ((2, 3, 1), (sz[0] - 1, sz[1] - 2, sz[2] - 3))
    mask = rng.random(tuple(h - l for l, h in zip(*box_z))) > 0.3
    ops, descr = [], []
    for (c0, c1, c2, c3) in OA.lorentz_coeffs(LORENTZ, DT):
        ops.append(AdeOp("Ez", 0, box_z[0], box_z[1], c0, c1, c2, c3, mask))
    descr.append(("lorentz", LORENTZ, "Ez", box_z, mask))
    box_x = ((0, 0, 0), sx)
    c = OA.drude_coeffs(1.2e16, 1e14, DT)
    ops.append(AdeOp("Ex", 1, box_x[0], box_x[1], c[0], c[1]))
    descr.append(("drude", (1.2e16, 1e14), "Ex", box_x, None))
    box_y = ((1, 2, 3), (sy[0] - 3, sy[1] - 1, sy[2] - 2))
    c = OA.debye_coeffs(2.0, 5.0, 3e-14, DT)
    ops.append(AdeOp("Ey", 2, box_y[0], box_y[1], c[0], c[1]))
    descr.append(("debye", (2.0, 5.0, 3e-14), "Ey", box_y, None))
    return ops, descr


def _seed(eng, rng):
    F = {c: rng.standard_normal(eng.field_shape(c)) * (1.0 if c[0] == "E" else 1 / 377.0) for c in COMPS}
    for c in COMPS:
        eng.upload(c, F[c])
    return F


def _states(eng, ids, ops):
    out = []
    for i, op in zip(ids, ops):
        out.append(eng.ade_state(i, 0))
        if op.kind == 0:
            out.append(eng.ade_state(i, 1))
    return out


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("het", [False, True], ids=["uniform", "het"])
@pytest.mark.parametrize("steps", [1, 5, 37])
def test_in_sweep_recursion_equals_post_kernel(dtype, het, steps):
    dims = (41, 37, 70)
    res = []
    for fused in (0, 1):
        rng = np.random.default_rng(11)
        eng, _ = _make(dims, dtype, het, rng)
        eng.set_option("ade_fused", fused)
        ops, _ = _ops(eng)
        ids = [eng.add_ade_op(op) for op in ops]
        sy = eng.field_shape("Ez")
        eng.add_source_op(pb.SourceOp("Ez", (7, 0, 0), (8, sy[1], sy[2]), 0))       # a source plane through the medium
        eng.set_tables(steps, rng.standard_normal((steps, 1)), np.zeros((steps, 0), dtype=np.complex128))
        _seed(eng, rng)
        l0 = eng.kernel_launches
        eng.run(steps)
        eng.sync()
        res.append(({c: eng.download(c) for c in COMPS}, _states(eng, ids, ops), eng.kernel_launches - l0))
        eng.close()
    (fa, sa, la), (fb, sb, lb) = res
    for c in COMPS:
        assert np.array_equal(fa[c], fb[c]), c
    assert len(sa) == len(sb) == 6
    for a, b in zip(sa, sb):
        assert (steps == 1 or np.abs(a).max() > 0) and np.array_equal(a, b)      # P_prev is still zero after one step
    if steps > 1:
        assert lb < la                                    # fewer launches: k_ade runs once per run / graph, not per step


@pytest.mark.parametrize("het", [False, True], ids=["uniform", "het"])
def test_coupled_mode_matches_its_oracle(het):
    dims, steps = (29, 33, 70), 6
    want = None
    for dtype in ("float64", "float32"):
        rng = np.random.default_rng(3)
        eng, coeffs = _make(dims, dtype, het, rng)
        eng.set_option("ade_coupled", 1)
        ops, descr = _ops(eng)
        ids = [eng.add_ade_op(op) for op in ops]
        F = _seed(eng, rng)
        eng.run(steps)
        got = {c: eng.download(c) for c in COMPS}
        gst = _states(eng, ids, ops)
        eng.close()
        if want is None:
            ades = []
            for kind, params, comp, (lo, hi), mask in descr:
                m = np.zeros(F[comp].shape)
                m[tuple(slice(l, h) for l, h in zip(lo, hi))] = 1.0 if mask is None else mask
                ades.append(OA.OAde(kind, params, DT, F[comp].shape, comp, m))
            for _ in range(steps):
                OA.coupled_step(F, coeffs, (D,) * 3, ades, DT, kernels)
            want = F
            wst = []
            for a, (kind, params, comp, (lo, hi), mask) in zip(ades, descr):
                box = tuple(slice(l, h) for l, h in zip(lo, hi))
                if kind == "lorentz":
                    for p, pp in zip(a.P, a.Pp):
                        wst += [p[box], pp[box]]
                else:
                    wst.append((a.J if kind == "drude" else a.P)[box])
            # the feedback must actually change E (otherwise this test proves nothing)
            plain = {c: v.copy() for c, v in _reseed(dims, het).items()}
            for _ in range(steps):
                kernels.step(plain, coeffs, (D,) * 3, False)
            assert not np.array_equal(plain["Ez"], want["Ez"])
        if dtype == "float64":
            for c in COMPS:
                assert np.array_equal(got[c], want[c]), (c, np.abs(got[c] - want[c]).max())
            for a, b in zip(gst, wst):
                assert np.array_equal(a, b)
        else:
            for c in COMPS:
                assert np.linalg.norm(got[c] - want[c]) <= 1e-4 * np.linalg.norm(want[c]), c


def _reseed(dims, het):
    rng = np.random.default_rng(3)
    if het:
        _eps(dims, rng)
    shapes = {"Ex": (0, 1, 1), "Ey": (1, 0, 1), "Ez": (1, 1, 0), "Hx": (1, 0, 0), "Hy": (0, 1, 0), "Hz": (0, 0, 1)}
    return {c: rng.standard_normal(tuple(d - s for d, s in zip(dims, shapes[c]))) * (1.0 if c[0] == "E" else 1 / 377.0)
            for c in COMPS}


def test_coupled_mode_needs_a_fused_sweep():
    from prismo_b200 import _lib

    eng = pb.Engine(3, (12, 12, 70), (D,) * 3, DT, flags=_lib.FLAG_TWO_PASS)
    eng.set_option("ade_coupled", 1)
    eng.add_ade_op(AdeOp("Ez", 1, (0, 0, 0), (3, 3, 3), 0.1, 0.9))
    with pytest.raises(RuntimeError, match="coupled dispersive media"):
        eng.run(1)
    eng.close()
