"""Load the UNMODIFIED reference (rithulkamesh/prismo) for oracle pinning.  Test infrastructure.

Two places can hold it: /root/reference/src (the build container) and oracle/_ref (a staged copy of the package made
by oracle/stage_reference.py: git-ignored, but it travels to the GPU box with gpurun like a built .so).  /root/reference
itself never exists on the GPU box; the ``-m gpu`` plugin tests use the staged copy and skip when it is absent.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_SRC = os.environ.get("PRISMO_REFERENCE_SRC") or (
    "/root/reference/src" if os.path.isdir("/root/reference/src/prismo") else _STAGED)


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "prismo"))


def install_shims() -> None:
    """Same two shims as oracle/_shim/sitecustomize.py, applied in-process."""
    sys.modules.setdefault("prismo.backends.metal_backend", None)
    if "polars" not in sys.modules:
        try:
            import polars  # noqa: F401
        except ImportError:
            pl = types.ModuleType("polars")
            pl.DataFrame = object
            sys.modules["polars"] = pl


def load():
    """Import and return the reference ``prismo`` package."""
    if "prismo" in sys.modules:
        return sys.modules["prismo"]
    if not available():
        raise ImportError(f"reference not present at {REF_SRC}")
    install_shims()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    return importlib.import_module("prismo")
