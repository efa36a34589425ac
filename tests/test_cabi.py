"""The C-ABI library loads and exports every symbol include/fdtd_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from prismo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fdtd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fdtd_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in fdtd_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes prototypes and header disagree"


def test_abi_version_and_struct_sizes():
    lib = _lib.load()
    assert lib.fdtd_abi_version() == _lib.ABI_VERSION
    for which, st in enumerate((_lib.Config, _lib.SourceOp, _lib.MonitorOp, _lib.AdeOp)):
        assert lib.fdtd_struct_size(which) == ctypes.sizeof(st)


def test_bad_arguments_are_reported_not_crashed():
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.Config(ndim=5, nx=8, ny=8, nz=8, dx=1, dy=1, dz=1, dt=1, dtype=0, device=0)
    assert lib.fdtd_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"ndim" in lib.fdtd_last_error()
    with pytest.raises(ValueError, match="ndim"):
        _lib.check(-1)
    assert lib.fdtd_run(None, 3) == -1 and lib.fdtd_destroy(None) == 0


def test_no_cpu_fallback_without_a_gpu():
    """On a box without CUDA the engine must fail loudly instead of stepping on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import prismo_b200 as pb

    with pytest.raises(RuntimeError, match="libfdtd_b200"):
        pb.Engine(3, (8, 8, 8), (1e-8, 1e-8, 1e-8), 1e-17)
    sim = pb.Simulation((0.5e-6, 0.5e-6, 0.5e-6), 20e6, pml_layers=2)
    with pytest.raises(RuntimeError):
        sim.step()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "prismo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "from tests" not in txt and "import tests" not in txt, f
