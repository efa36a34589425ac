"""Row a23 — anisotropic tensor update at function level (AnisotropicUpdater.update_e_from_curl_h /
update_h_from_curl_e, /root/reference/src/prismo/materials/tensor.py:482-588).

tests/golden/aniso.npz holds the REAL reference's outputs (tests/golden/make_golden.py) for the cases of
tests/tensor_cases.py with float64 and float32 inputs: diagonal scalar, diagonal per-cell, full uniform (rotated
uniaxial), full per-cell.  Bar: bit-exact values AND NumPy's result dtypes.
  not gpu : the oracle restatement; the host mirror's promotion / staging logic against a NumPy stand-in of the kernel
  gpu     : the CUDA kernel through the C ABI (fdtd_tensor_update), incl. the chunked path
"""
import os

import numpy as np
import pytest

from tests import tensor_cases as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aniso.npz")


def _material(pb, spec):
    return pb.TensorMaterial(pb.TensorComponents(**spec["eps"]),
                             None if spec["mu"] is None else pb.TensorComponents(**spec["mu"]))


def _check_against_golden(update_pair):
    """update_pair(name, spec, f, c) -> (E triple, H triple)."""
    z = np.load(GOLDEN)
    seen = 0
    for name, spec in T.cases().items():
        for dt in ("float64", "float32"):
            f, c = T.inputs(dt)
            e, h = update_pair(name, spec, f, c)
            for which, arrs in (("e", e), ("h", h)):
                for k, a in enumerate(arrs):
                    g = z[f"{name}.{dt}.{which}{k}"]
                    assert a.dtype == g.dtype, f"{name} {dt} {which}{k}: dtype {a.dtype} != {g.dtype}"
                    assert a.shape == g.shape
                    assert np.array_equal(a, g), f"{name} {dt} {which}{k}: max |diff| {np.abs(a - g).max()}"
                    seen += 1
    assert seen == len(z.files) == 48


def test_oracle_matches_reference_golden():
    import prismo_b200 as pb
    from oracle import tensor as OT

    def run(name, spec, f, c):
        m = _material(pb, spec)
        if m.is_diagonal:
            return (OT.update_e(f, c, T.DT, diagonal=(m.epsilon.xx, m.epsilon.yy, m.epsilon.zz)),
                    OT.update_h(f, c, T.DT, diagonal=(m.mu.xx, m.mu.yy, m.mu.zz)))
        return (OT.update_e(f, c, T.DT, inverse=m.get_inverse_epsilon()), OT.update_h(f, c, T.DT, inverse=m.get_inverse_mu()))

    _check_against_golden(run)


def test_mirror_promotion_and_staging_on_kernel_stand_in(monkeypatch):
    import prismo_b200 as pb
    from prismo_b200 import _lib
    from tests._fake_engine import FakeTensorLib

    fake = FakeTensorLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)

    def run(name, spec, f, c):
        upd = pb.AnisotropicUpdater(_material(pb, spec), T.DT)
        return upd.update_e_from_curl_h(tuple(f), tuple(c)), upd.update_h_from_curl_e(tuple(f), tuple(c))

    _check_against_golden(run)
    modes = {(d, m) for d, _, _, m in fake.calls}
    # (dtype, mode): all-float32 diagonal; float64 diagonal; full tensor; float32 curl against a float64 tensor array
    # (product rounded in float32, then float64)
    assert modes == {(0, 0), (1, 0), (1, 1), (1, 2)}


def test_mixed_precision_stages(monkeypatch):
    """float64 field + float32 curl + Python-float entry: NumPy multiplies and divides in float32, adds in float64."""
    import prismo_b200 as pb
    from prismo_b200 import _lib
    from tests._fake_engine import FakeTensorLib

    fake = FakeTensorLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    rng = np.random.default_rng(1)
    f = [rng.standard_normal(50) for _ in range(3)]
    c = [(rng.standard_normal(50) * 1e6).astype(np.float32) for _ in range(3)]
    s = T.DT / 8.854187817e-12
    got = pb.tensor_update(f, c, s, (2.25, np.float64(4.0), np.full(50, 3.0, dtype=np.float32)))
    want = [f[0] + s * c[0] / 2.25, f[1] + s * c[1] / np.float64(4.0), f[2] + s * c[2] / np.full(50, 3.0, dtype=np.float32)]
    for a, b in zip(got, want):
        assert a.dtype == b.dtype == np.float64 and np.array_equal(a, b)
    assert sorted(m for _, _, _, m in fake.calls) == [2, 6]      # np.float64 entry: only the product is float32


def test_updater_rejects_bad_backend_like_the_reference():
    import prismo_b200 as pb

    with pytest.raises(TypeError):
        pb.AnisotropicUpdater(pb.TensorMaterial(pb.TensorComponents(xx=2.0, yy=2.0, zz=2.0)), 1e-17, backend=3)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [None, "37"])
def test_cuda_kernel_matches_reference_golden(chunk, monkeypatch):
    import prismo_b200 as pb

    if chunk:
        monkeypatch.setenv("FDTD_B200_TENSOR_CHUNK", chunk)

    def run(name, spec, f, c):
        upd = pb.AnisotropicUpdater(_material(pb, spec), T.DT)
        return upd.update_e_from_curl_h(tuple(f), tuple(c)), upd.update_h_from_curl_e(tuple(f), tuple(c))

    _check_against_golden(run)


@pytest.mark.gpu
def test_cuda_kernel_mixed_precision_and_large():
    import prismo_b200 as pb

    rng = np.random.default_rng(1)
    n = 300_001
    f = [rng.standard_normal(n) for _ in range(3)]
    c = [(rng.standard_normal(n) * 1e6).astype(np.float32) for _ in range(3)]
    s = T.DT / 8.854187817e-12
    d = (2.25, np.float64(4.0), np.full(n, 3.0, dtype=np.float32))
    got = pb.tensor_update(f, c, s, d, negative=True)
    want = [f[k] - s * c[k] / d[k] for k in range(3)]
    for a, b in zip(got, want):
        assert a.dtype == b.dtype == np.float64 and np.array_equal(a, b)


@pytest.mark.reference
def test_plugin_routes_reference_anisotropic_updater(monkeypatch, ref):
    """prismo.set_backend('b200') + the reference's own TensorMaterial / AnisotropicUpdater objects: the two update
    methods run through fdtd_tensor_update (stand-in here) and reproduce the NumPy backend bit for bit."""
    import prismo_b200 as pb
    from prismo_b200 import _lib
    from prismo.materials.tensor import AnisotropicUpdater, TensorComponents, TensorMaterial
    from tests._fake_engine import FakeTensorLib

    real_load = _lib.load
    fake = FakeTensorLib()
    pb.register()
    try:
        ref.set_backend("b200")                            # loads the real library once (fails loudly if missing)
        monkeypatch.setattr(_lib, "load", lambda: fake)

        def run(name, spec, f, c):
            mat = TensorMaterial(TensorComponents(**spec["eps"]), None if spec["mu"] is None else TensorComponents(**spec["mu"]))
            upd = AnisotropicUpdater(mat, T.DT)
            assert upd.backend.name == "b200"
            return upd.update_e_from_curl_h(tuple(f), tuple(c)), upd.update_h_from_curl_e(tuple(f), tuple(c))

        _check_against_golden(run)
        assert len(fake.calls) >= 16
        # an updater bound to the NumPy backend is left alone
        n_calls = len(fake.calls)
        mat = TensorMaterial(TensorComponents(xx=2.0, yy=3.0, zz=4.0), backend="numpy")
        upd = AnisotropicUpdater(mat, T.DT, backend="numpy")
        f, c = T.inputs("float64")
        upd.update_e_from_curl_h(tuple(f), tuple(c))
        assert len(fake.calls) == n_calls
    finally:
        monkeypatch.setattr(_lib, "load", real_load)
        ref.set_backend("numpy")


def test_c_entry_argument_checks_without_a_gpu():
    """fdtd_tensor_update validates its pointer tables before touching CUDA: exercised through the real library."""
    import ctypes as C

    from prismo_b200 import _lib

    lib = _lib.load()
    a = np.zeros(4)
    tab = lambda *ptrs: (C.c_void_p * 3)(*ptrs)          # noqa: E731
    coef = (C.c_double * 9)(*([1.0] * 9))
    none9 = (C.c_void_p * 9)()
    p = a.ctypes.data
    # nothing to do: n == 0, or no field component selected
    assert lib.fdtd_tensor_update(0, _lib.F64, 0, tab(p, p, p), tab(p, p, p), tab(p, p, p), 1.0, 0, 0, coef, none9) == 0
    assert lib.fdtd_tensor_update(0, _lib.F64, 4, tab(None, None, None), tab(p, p, p), tab(p, p, p), 1.0, 0, 0, coef, none9) == 0
    # a selected component without output / without its curl; the full tensor needs all three curls; bad dtype
    for args in ((0, _lib.F64, 4, tab(p, None, None), tab(p, None, None), tab(None, None, None), 1.0, 0, 0, coef, none9),
                 (0, _lib.F64, 4, tab(p, None, None), tab(None, p, p), tab(p, None, None), 1.0, 0, 0, coef, none9),
                 (0, _lib.F64, 4, tab(p, None, None), tab(p, p, None), tab(p, None, None), 1.0, 0, 1, coef, none9),
                 (0, 7, 4, tab(p, p, p), tab(p, p, p), tab(p, p, p), 1.0, 0, 0, coef, none9)):
        with pytest.raises(ValueError):
            _lib.check(lib.fdtd_tensor_update(*args))
