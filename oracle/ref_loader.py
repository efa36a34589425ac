"""Load the UNMODIFIED reference (rithulkamesh/prismo) for oracle pinning.  Test infrastructure.

Works only where /root/reference exists (the build container).  Nothing that runs on the GPU
box (``-m gpu`` tests, smoke(), bench.py) may call this; they use the committed goldens instead.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_SRC = os.environ.get("PRISMO_REFERENCE_SRC", "/root/reference/src")


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
