"""ctypes binding of libfdtd_b200.so — the C ABI declared in include/fdtd_b200.h.

There is no fallback: if the shared library is missing or cannot be loaded, importing the engine
raises.  Build it with ``python __graft_entry__.py build`` (or ``prismo_b200/csrc/build.sh``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfdtd_b200.so")

ABI_VERSION = 1
F32, F64 = 0, 1
COMPONENTS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
COMP_ID = {c: i for i, c in enumerate(COMPONENTS)}
FLAG_NO_GRAPH, FLAG_TWO_PASS, FLAG_YEE, FLAG_FAST_F64 = 1, 2, 4, 8

_ERR_TYPES = {-1: ValueError, -2: RuntimeError, -3: MemoryError, -4: RuntimeError}


class Config(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
        ("dtype", C.c_int32), ("device", C.c_int32), ("nx_global", C.c_int32),
        ("x_offset", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32),
    ]


class SourceOp(C.Structure):
    _fields_ = [
        ("component", C.c_int32), ("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3), ("table", C.c_int32),
        ("profile", C.POINTER(C.c_double)), ("divisor", C.c_double), ("group", C.c_int32),
        ("reserved", C.c_int32),
    ]


class MonitorOp(C.Structure):
    _fields_ = [
        ("component", C.c_int32), ("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3), ("record", C.c_int32),
        ("n_freq", C.c_int32), ("phasor_col", C.c_int32), ("reserved", C.c_int32),
    ]


class AdeOp(C.Structure):
    _fields_ = [
        ("component", C.c_int32), ("kind", C.c_int32), ("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3),
        ("c0", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("c3", C.c_double),
        ("mask", C.POINTER(C.c_uint8)),
    ]


_P = C.c_void_p
_PROTOS = {
    # name: (restype, argtypes)
    "fdtd_abi_version": (C.c_int, []),
    "fdtd_struct_size": (C.c_int, [C.c_int32]),
    "fdtd_last_error": (C.c_char_p, []),
    "fdtd_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "fdtd_destroy": (C.c_int, [_P]),
    "fdtd_set_uniform_coeffs": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double]),
    "fdtd_set_coeffs": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32]),
    "fdtd_set_coeffs_aniso": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32]),
    "fdtd_rasterize": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, _P, _P, _P, C.c_int32, _P]),
    "fdtd_download_coeffs": (C.c_int, [_P, C.c_int32, _P, C.c_int32]),
    "fdtd_set_cpml": (C.c_int, [_P, C.c_int32, _P]),
    "fdtd_upload_field": (C.c_int, [_P, C.c_int32, _P, C.c_int32]),
    "fdtd_download_field": (C.c_int, [_P, C.c_int32, _P, C.c_int32]),
    "fdtd_zero_fields": (C.c_int, [_P]),
    "fdtd_field_device_ptr": (C.c_int, [_P, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fdtd_clear_ops": (C.c_int, [_P]),
    "fdtd_add_source_op": (C.c_int, [_P, C.POINTER(SourceOp)]),
    "fdtd_add_monitor_op": (C.c_int, [_P, C.POINTER(MonitorOp), C.POINTER(C.c_int32)]),
    "fdtd_add_flux_op": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "fdtd_download_flux": (C.c_int, [_P, C.c_int32, _P, C.c_int32]),
    "fdtd_add_ade_op": (C.c_int, [_P, C.POINTER(AdeOp), C.POINTER(C.c_int32)]),
    "fdtd_download_ade": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "fdtd_upload_ade": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "fdtd_set_tables": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P]),
    "fdtd_run": (C.c_int, [_P, C.c_int32]),
    "fdtd_update_h": (C.c_int, [_P]),
    "fdtd_update_e": (C.c_int, [_P]),
    "fdtd_sync": (C.c_int, [_P]),
    "fdtd_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "fdtd_timer_start": (C.c_int, [_P]),
    "fdtd_timer_stop": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "fdtd_run_profiled": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double)]),
    "fdtd_pass": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "fdtd_sweep": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "fdtd_post_step": (C.c_int, [_P, _P]),
    "fdtd_ipc_export": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "fdtd_ipc_connect": (C.c_int, [_P, _P, C.c_int32]),
    "fdtd_slab_run": (C.c_int, [_P, C.c_int32]),
    "fdtd_slab_sync": (C.c_int, [_P]),
    "fdtd_halo_ptrs": (C.c_int, [_P, C.c_int32, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64)]),
    "fdtd_download_records": (C.c_int, [_P, C.c_int32, _P, C.c_int32]),
    "fdtd_download_dft": (C.c_int, [_P, C.c_int32, _P]),
    "fdtd_upload_dft": (C.c_int, [_P, C.c_int32, _P]),
    "fdtd_download_box": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "fdtd_field_checksum": (C.c_int, [_P, C.c_int32, _P, C.c_int32]),
    "fdtd_mode_overlap": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int32, _P, _P]),
    "fdtd_steps_done": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "fdtd_kernel_launches": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "fdtd_mem_info": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fdtd_device_mem_info": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fdtd_device_sync": (C.c_int, [C.c_int32]),
    "fdtd_host_alloc": (C.c_int, [C.c_int64, C.POINTER(_P)]),
    "fdtd_host_free": (C.c_int, [_P]),
    "fdtd_tensor_update": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                                     C.POINTER(C.c_void_p)]),
    "fdtd_plan_segments": (C.c_int, [C.c_int32, C.POINTER(C.c_uint8), C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises OSError if it was not built: no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(
            f"{LIB_PATH} not found: the B200 engine has no CPU fallback. "
            "Build it with `python __graft_entry__.py build`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    got = lib.fdtd_abi_version()
    if got != ABI_VERSION:
        raise OSError(f"{LIB_PATH}: ABI version {got}, binding expects {ABI_VERSION}")
    from .geometry import ShapeStruct

    for which, st in enumerate((Config, SourceOp, MonitorOp, AdeOp, ShapeStruct)):
        if lib.fdtd_struct_size(which) != C.sizeof(st):
            raise OSError(f"{LIB_PATH}: struct {st.__name__} is {lib.fdtd_struct_size(which)} bytes in C, "
                          f"{C.sizeof(st)} in the binding")
    _lib = lib
    return lib


def pinned_empty(shape, dtype):
    """NumPy array over page-locked host memory (cudaHostAlloc), freed when the array's buffer is collected.
    Falls back to pageable memory for tiny arrays (not worth a driver call)."""
    import weakref

    import numpy as np

    dt = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(int(v) for v in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    if nbytes < (1 << 16):
        return np.empty(shape, dtype=dt)
    lib = load()
    p = _P()
    if lib.fdtd_host_alloc(nbytes, C.byref(p)) != 0:      # no driver / pinning refused: a pageable mirror still works
        return np.empty(shape, dtype=dt)
    buf = (C.c_char * nbytes).from_address(p.value)
    weakref.finalize(buf, lib.fdtd_host_free, p.value)
    return np.frombuffer(buf, dtype=dt).reshape(shape)


def check(rc: int) -> None:
    """Map a C status to the Python exception type the reference would raise (SURVEY §8b errors)."""
    if rc == 0:
        return
    msg = load().fdtd_last_error().decode("utf-8", "replace")
    raise _ERR_TYPES.get(rc, RuntimeError)(f"libfdtd_b200: {msg}")
