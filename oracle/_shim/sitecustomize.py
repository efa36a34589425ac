"""Import shims that let the UNMODIFIED reference (rithulkamesh/prismo) import on Linux.

Test infrastructure only.  Put this directory first on PYTHONPATH/sys.path, followed by
/root/reference/src.  Two defects of the reference are papered over:
  * backends/metal_backend.py:85 evaluates an annotation naming MTLBuffer when PyObjC-Metal is
    absent (NameError); upstream only catches ImportError (backend_manager.py:22-26).  Mapping
    the module to None turns the failure into the ImportError the reference already handles.
  * io/exporters/parquet_exporter.py:301 annotates with polars.DataFrame; polars is absent.
"""
import sys
import types

sys.modules.setdefault("prismo.backends.metal_backend", None)
try:  # pragma: no cover - depends on the image
    import polars  # noqa: F401
except ImportError:
    _pl = types.ModuleType("polars")
    _pl.DataFrame = object
    sys.modules["polars"] = _pl
