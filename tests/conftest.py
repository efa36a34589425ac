import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def _cuda_devices() -> int:
    """Number of usable CUDA devices, asked of the CUDA runtime directly (0 if the library or driver is missing)."""
    import ctypes

    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU box must not fail the GPU tests: skip them when there is no device or no library.
    (`-m gpu` on the B200 box runs them all; the product itself still raises without its CUDA library.)"""
    so = os.path.join(ROOT, "prismo_b200", "libfdtd_b200.so")
    n = _cuda_devices() if os.path.exists(so) else 0
    if n > 0:
        return
    why = "no CUDA device" if os.path.exists(so) else "libfdtd_b200.so not built"
    skip = pytest.mark.skip(reason=f"GPU test: {why}")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ref():
    """The real reference package, or skip where it does not exist (the GPU box)."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present")
    return ref_loader.load()
