"""``Engine`` — thin object wrapper over the C ABI (include/fdtd_b200.h).  One engine = one GPU.

All arithmetic happens in libfdtd_b200.so; this file only marshals NumPy buffers and op descriptors.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import COMP_ID, COMPONENTS, F32, F64


@dataclass
class SourceOp:
    """F[box] += amp_table[step, table] (* profile / divisor).  Box in the component's LOCAL index space."""
    component: str
    lo: tuple
    hi: tuple
    table: int
    profile: Optional[np.ndarray] = None
    divisor: float = 1.0
    group: int = 0
    ghost: bool = False     # x-slabs: the right neighbour's op on our ghost planes [nx, nx+3) (two-step sweep only)


@dataclass
class MonitorOp:
    """Sample F[box] every step: keep it (record) and/or add it to a running DFT over ``n_freq`` phasors."""
    component: str
    lo: tuple
    hi: tuple
    record: bool = False
    n_freq: int = 0
    phasor_col: int = 0
    shape: tuple = field(default=(), compare=False)     # box shape in the array's own rank (2 or 3 axes)


@dataclass
class AdeOp:
    """Cell-local ADE recursion driven by E[component] after every step (see fdtd_ade_op)."""
    component: str
    kind: int                      # 0 Lorentz, 1 Drude, 2 Debye
    lo: tuple
    hi: tuple
    c0: float
    c1: float
    c2: float = 0.0
    c3: float = 0.0
    mask: Optional[np.ndarray] = None
    shape: tuple = field(default=(), compare=False)


def _dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise TypeError(f"unsupported dtype {dt}; the engine stores float32 or float64")


def _pad3(t, fill):
    t = tuple(int(v) for v in t)
    return t + (fill,) * (3 - len(t))


class Engine:
    def __init__(self, ndim: int, dims: Sequence[int], spacing: Sequence[float], dt: float, dtype="float64",
                 device: int = 0, nx_global: Optional[int] = None, x_offset: int = 0, flags: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.ndim = int(ndim)
        self.dims = tuple(int(d) for d in dims)
        self.dtype = np.dtype(dtype)
        self.dt = float(dt)
        self.nx_global = int(nx_global) if nx_global else self.dims[0]
        self.x_offset = int(x_offset)
        cfg = _lib.Config(ndim=self.ndim, nx=self.dims[0], ny=self.dims[1], nz=self.dims[2] if self.ndim == 3 else 1,
                          dx=spacing[0], dy=spacing[1], dz=spacing[2] if self.ndim == 3 else 0.0, dt=self.dt,
                          dtype=_dtype_code(self.dtype), device=int(device), nx_global=self.nx_global,
                          x_offset=self.x_offset, flags=int(flags), reserved=0)
        _lib.check(self._lib.fdtd_create(C.byref(cfg), C.byref(self._h)))
        self._mon_ops: list = []
        self._ade_ops: list = []

    # ---- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fdtd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- shapes -------------------------------------------------------------------------------------
    def field_shape(self, comp: str) -> tuple:
        """LOCAL host shape of a component (core/grid.py:157-168; the last x-slab owns the short end)."""
        from .grid import SHORT_AXES

        n = list(self.dims[: self.ndim])
        last = self.x_offset + self.dims[0] == self.nx_global
        for ax in SHORT_AXES[comp]:
            if ax >= self.ndim:
                continue
            if ax == 0 and not last:
                continue
            n[ax] -= 1
        return tuple(n)

    # ---- coefficients ---------------------------------------------------------------------------------
    def set_uniform_coeffs(self, ca, cb, da, db):
        _lib.check(self._lib.fdtd_set_uniform_coeffs(self._h, float(ca), float(cb), float(da), float(db)))

    def set_coeffs(self, Ca, Cb, Da, Db):
        """Cell-centred arrays of shape (planes, ny, nz) with planes = nx or nx+1 (right-neighbour ghost)."""
        arrs = []
        planes = None
        for a in (Ca, Cb, Da, Db):
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.ndim == 2:
                a = a[:, :, None]
            want = self.dims[1:] if self.ndim == 3 else (self.dims[1], 1)
            if a.shape[1:] != tuple(want) or a.shape[0] not in (self.dims[0], self.dims[0] + 1):
                raise ValueError(f"coefficient array shape {a.shape} does not match grid {self.dims}")
            planes = a.shape[0] if planes is None else planes
            if a.shape[0] != planes:
                raise ValueError("coefficient arrays disagree on the number of planes")
            arrs.append(a)
        _lib.check(self._lib.fdtd_set_coeffs(self._h, *[a.ctypes.data_as(C.c_void_p) for a in arrs], int(planes)))

    def set_coeffs_aniso(self, Ca, Cbx, Cby, Cbz, Da, Db):
        """OPT-IN extension: per-component Cb (diagonal permittivity tensor) inside the E stage; 3-D, parity sweeps."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Ca, Cbx, Cby, Cbz, Da, Db)]
        planes = arrs[0].shape[0]
        for a in arrs:
            if a.ndim != 3 or a.shape[1:] != tuple(self.dims[1:]) or a.shape[0] != planes or \
                    planes not in (self.dims[0], self.dims[0] + 1):
                raise ValueError(f"coefficient array shape {a.shape} does not match grid {self.dims}")
        _lib.check(self._lib.fdtd_set_coeffs_aniso(self._h, *[a.ctypes.data_as(C.c_void_p) for a in arrs], int(planes)))

    def rasterize(self, shapes, x, y, z=None, background=None) -> None:
        """Shape list -> Ca, Cb, Da, Db on the device (geometry/shapes.py:71-99 + core/solver.py:113-133).  x: nx (or nx+1
        on a slab with a right neighbour) cell coordinates; y, z: ny, nz.  Later shapes paint over earlier ones."""
        from . import geometry

        arr, n, verts = geometry.lower_shapes(shapes)
        bg = geometry.background_values(background)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if self.ndim == 3:
            if z is None:
                raise ValueError("a 3-D grid needs z coordinates")
            z = np.ascontiguousarray(z, dtype=np.float64)
            if z.size != self.dims[2]:
                raise ValueError(f"z has {z.size} entries, grid has {self.dims[2]}")
        else:
            z = None
        if y.size != self.dims[1]:
            raise ValueError(f"y has {y.size} entries, grid has {self.dims[1]}")
        _lib.check(self._lib.fdtd_rasterize(
            self._h, arr, n, verts.ctypes.data_as(C.c_void_p) if len(verts) else None, len(verts),
            x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
            z.ctypes.data_as(C.c_void_p) if z is not None else None, int(x.size), bg.ctypes.data_as(C.c_void_p)))

    def download_coeffs(self, which, planes: Optional[int] = None) -> np.ndarray:
        """Device coefficient array as host fp64: which = 'Ca' | 'Cb' | 'Da' | 'Db' | 'Cby' | 'Cbz' (or 0..5)."""
        idx = {"Ca": 0, "Cb": 1, "Cbx": 1, "Da": 2, "Db": 3, "Cby": 4, "Cbz": 5}.get(which, which)
        planes = self.dims[0] if planes is None else int(planes)
        shape = (planes, self.dims[1], self.dims[2]) if self.ndim == 3 else (planes, self.dims[1])
        out = np.empty(shape, dtype=np.float64)
        _lib.check(self._lib.fdtd_download_coeffs(self._h, int(idx), out.ctypes.data_as(C.c_void_p), planes))
        return out

    def set_cpml(self, thickness: int, coef: Optional[np.ndarray]) -> None:
        """Physics mode only.  coef: concatenation over x, y, z of (6, N_axis) arrays (see cpml.py)."""
        if thickness == 0:
            _lib.check(self._lib.fdtd_set_cpml(self._h, 0, None))
            return
        c = np.ascontiguousarray(coef, dtype=np.float64)
        if c.size != 6 * sum(self.dims):
            raise ValueError(f"CPML coefficient table has {c.size} entries, expected {6 * sum(self.dims)}")
        _lib.check(self._lib.fdtd_set_cpml(self._h, int(thickness), c.ctypes.data_as(C.c_void_p)))

    # ---- fields -------------------------------------------------------------------------------------------
    def upload(self, comp: str, array) -> None:
        a = np.asarray(array)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        a = np.ascontiguousarray(a)
        if a.shape != self.field_shape(comp):
            raise ValueError(f"Shape mismatch for {comp}: expected {self.field_shape(comp)}, got {a.shape}")
        _lib.check(self._lib.fdtd_upload_field(self._h, COMP_ID[comp], a.ctypes.data_as(C.c_void_p), _dtype_code(a.dtype)))

    def download(self, comp: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        shape = self.field_shape(comp)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        if out.shape != shape or not out.flags.c_contiguous or out.dtype not in (np.float32, np.float64):
            tmp = np.empty(shape, dtype=np.float64)
            self.download(comp, tmp)
            out[...] = tmp
            return out
        _lib.check(self._lib.fdtd_download_field(self._h, COMP_ID[comp], out.ctypes.data_as(C.c_void_p), _dtype_code(out.dtype)))
        return out

    def download_box(self, comp: str, lo, hi) -> np.ndarray:
        """fields[comp][lo:hi] as host fp64 without moving the whole array (fdtd_download_box)."""
        lo3, hi3 = _pad3(lo, 0), _pad3(hi, 1)
        out = np.empty(tuple(h - l for l, h in zip(lo, hi)), dtype=np.float64)
        _lib.check(self._lib.fdtd_download_box(self._h, COMP_ID[comp], (C.c_int32 * 3)(*lo3), (C.c_int32 * 3)(*hi3),
                                               out.ctypes.data_as(C.c_void_p)))
        return out

    def plane_checksums(self, comp: str) -> np.ndarray:
        """(planes, 2) uint64: order- and decomposition-independent checksums of each x-plane (fdtd_field_checksum)."""
        n = self.field_shape(comp)[0]
        out = np.zeros((n, 2), dtype=np.uint64)
        _lib.check(self._lib.fdtd_field_checksum(self._h, COMP_ID[comp], out.ctypes.data_as(C.c_void_p), n))
        return out

    def zero_fields(self):
        _lib.check(self._lib.fdtd_zero_fields(self._h))

    def device_ptr(self, comp: str):
        p, ps, rs, pl = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self._lib.fdtd_field_device_ptr(self._h, COMP_ID[comp], C.byref(p), C.byref(ps), C.byref(rs), C.byref(pl)))
        return p.value, ps.value, rs.value, pl.value

    # ---- ops --------------------------------------------------------------------------------------------------
    def clear_ops(self):
        self.ops_epoch = getattr(self, "ops_epoch", 0) + 1      # device handles to monitor ops die here
        _lib.check(self._lib.fdtd_clear_ops(self._h))
        self._mon_ops = []
        self._ade_ops = []

    def add_source_op(self, op: SourceOp) -> None:
        c = _lib.SourceOp()
        c.component = COMP_ID[op.component]
        c.lo[:] = _pad3(op.lo, 0)
        c.hi[:] = _pad3(op.hi, 1)
        c.table, c.divisor, c.group, c.reserved = int(op.table), float(op.divisor), int(op.group), int(bool(op.ghost))
        prof = None
        if op.profile is not None:
            prof = np.ascontiguousarray(op.profile, dtype=np.float64)
            want = tuple(h - l for l, h in zip(op.lo, op.hi))
            if prof.shape != want:
                raise ValueError(f"profile shape {prof.shape} != box shape {want}")
            c.profile = prof.ctypes.data_as(C.POINTER(C.c_double))
        _lib.check(self._lib.fdtd_add_source_op(self._h, C.byref(c)))

    def add_monitor_op(self, op: MonitorOp) -> int:
        c = _lib.MonitorOp()
        c.component = COMP_ID[op.component]
        c.lo[:] = _pad3(op.lo, 0)
        c.hi[:] = _pad3(op.hi, 1)
        c.record, c.n_freq, c.phasor_col, c.reserved = int(bool(op.record)), int(op.n_freq), int(op.phasor_col), 0
        mid = C.c_int32()
        _lib.check(self._lib.fdtd_add_monitor_op(self._h, C.byref(c), C.byref(mid)))
        op.shape = tuple(h - l for l, h in zip(op.lo, op.hi))
        self._mon_ops.append(op)
        return mid.value

    def add_flux_op(self, direction: str, lo, hi) -> int:
        """Region-correct power sum((E x H)_n) through box [lo, hi) (valid for all six components), every step."""
        a, b = (C.c_int32 * 3)(*_pad3(lo, 0)), (C.c_int32 * 3)(*_pad3(hi, 1))
        fid = C.c_int32()
        _lib.check(self._lib.fdtd_add_flux_op(self._h, "xyz".index(direction), a, b, C.byref(fid)))
        return fid.value

    def flux(self, flux_id: int, steps: int) -> np.ndarray:
        out = np.zeros(steps, dtype=np.float64)
        if steps:
            _lib.check(self._lib.fdtd_download_flux(self._h, flux_id, out.ctypes.data_as(C.c_void_p), int(steps)))
        return out

    def add_ade_op(self, op: AdeOp) -> int:
        c = _lib.AdeOp()
        c.component, c.kind = COMP_ID[op.component], int(op.kind)
        c.lo[:] = _pad3(op.lo, 0)
        c.hi[:] = _pad3(op.hi, 1)
        c.c0, c.c1, c.c2, c.c3 = float(op.c0), float(op.c1), float(op.c2), float(op.c3)
        op.shape = tuple(h - l for l, h in zip(op.lo, op.hi))
        m = None
        if op.mask is not None:
            m = np.ascontiguousarray(np.asarray(op.mask) != 0, dtype=np.uint8)
            if m.shape != op.shape:
                raise ValueError(f"mask shape {m.shape} != box shape {op.shape}")
            c.mask = m.ctypes.data_as(C.POINTER(C.c_uint8))
        aid = C.c_int32()
        _lib.check(self._lib.fdtd_add_ade_op(self._h, C.byref(c), C.byref(aid)))
        self._ade_ops.append(op)
        return aid.value

    def ade_state(self, ade_id: int, which: int = 0) -> np.ndarray:
        out = np.zeros(self._ade_ops[ade_id].shape, dtype=np.float64)
        if out.size:
            _lib.check(self._lib.fdtd_download_ade(self._h, ade_id, which, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_ade_state(self, ade_id: int, which: int, values) -> None:
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.shape != self._ade_ops[ade_id].shape:
            raise ValueError(f"ADE state shape {v.shape} != {self._ade_ops[ade_id].shape}")
        if v.size:
            _lib.check(self._lib.fdtd_upload_ade(self._h, ade_id, which, v.ctypes.data_as(C.c_void_p)))

    def set_tables(self, n_steps: int, amp: Optional[np.ndarray] = None, phasors: Optional[np.ndarray] = None):
        """amp: (n_steps, n_amp) float64; phasors: (n_steps, n_phasor) complex128."""
        amp = np.zeros((n_steps, 0)) if amp is None else np.ascontiguousarray(amp, dtype=np.float64).reshape(n_steps, -1)
        ph = (np.zeros((n_steps, 0), dtype=np.complex128) if phasors is None
              else np.ascontiguousarray(phasors, dtype=np.complex128).reshape(n_steps, -1))
        _lib.check(self._lib.fdtd_set_tables(self._h, int(n_steps), amp.shape[1], amp.ctypes.data_as(C.c_void_p),
                                             ph.shape[1], ph.ctypes.data_as(C.c_void_p)))

    # ---- stepping ---------------------------------------------------------------------------------------------------
    def run(self, n_steps: int):
        _lib.check(self._lib.fdtd_run(self._h, int(n_steps)))

    def update_h(self):
        _lib.check(self._lib.fdtd_update_h(self._h))

    def update_e(self):
        _lib.check(self._lib.fdtd_update_e(self._h))

    def sync(self):
        _lib.check(self._lib.fdtd_sync(self._h))

    def set_option(self, key: str, value: int) -> None:
        _lib.check(self._lib.fdtd_set_option(self._h, key.encode(), int(value)))

    def timer_start(self):
        _lib.check(self._lib.fdtd_timer_start(self._h))

    def timer_stop(self) -> float:
        """Milliseconds between timer_start and now, from CUDA events on the engine's stream (blocks)."""
        ms = C.c_double()
        _lib.check(self._lib.fdtd_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def run_profiled(self, n_steps: int) -> dict:
        """n real steps with events between kernels; summed kernel times in ms."""
        out = (C.c_double * 4)()
        _lib.check(self._lib.fdtd_run_profiled(self._h, int(n_steps), out))
        return {"h_or_fused_ms": out[0], "e_ms": out[1], "post_ms": out[2], "total_ms": out[3]}

    def run_pass(self, phase: int, part: int = 2, stream: int = 0):
        _lib.check(self._lib.fdtd_pass(self._h, int(phase), int(part), C.c_void_p(stream)))

    def sweep(self, i_begin: int, i_end: int, flip: bool, stream: int = 0):
        """Fused step over local planes [i_begin, i_end) (see fdtd_sweep in include/fdtd_b200.h)."""
        _lib.check(self._lib.fdtd_sweep(self._h, int(i_begin), int(i_end), int(bool(flip)), C.c_void_p(stream)))

    def ipc_export(self) -> bytes:
        n = C.c_int32(0)
        _lib.check(self._lib.fdtd_ipc_export(self._h, None, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _lib.check(self._lib.fdtd_ipc_export(self._h, buf, C.byref(n)))
        return buf.raw[: n.value]

    def ipc_connect(self, left_blob: Optional[bytes], has_right: bool) -> None:
        buf = C.create_string_buffer(left_blob, len(left_blob)) if left_blob else None
        _lib.check(self._lib.fdtd_ipc_connect(self._h, buf, int(bool(has_right))))

    def slab_run(self, n_steps: int) -> None:
        _lib.check(self._lib.fdtd_slab_run(self._h, int(n_steps)))

    def slab_sync(self) -> None:
        _lib.check(self._lib.fdtd_slab_sync(self._h))

    def post_step(self, stream: int = 0):
        _lib.check(self._lib.fdtd_post_step(self._h, C.c_void_p(stream)))

    def halo_ptrs(self, comp: str):
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self._lib.fdtd_halo_ptrs(self._h, COMP_ID[comp], C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    # ---- read-out ------------------------------------------------------------------------------------------------------
    def records(self, monitor_id: int, steps: int) -> np.ndarray:
        op = self._mon_ops[monitor_id]
        out = np.empty((steps,) + op.shape, dtype=np.float64)
        if out.size:
            _lib.check(self._lib.fdtd_download_records(self._h, monitor_id, out.ctypes.data_as(C.c_void_p), int(steps)))
        return out

    def dft(self, monitor_id: int) -> np.ndarray:
        op = self._mon_ops[monitor_id]
        out = np.zeros((op.n_freq,) + op.shape, dtype=np.complex128)
        if out.size:
            _lib.check(self._lib.fdtd_download_dft(self._h, monitor_id, out.ctypes.data_as(C.c_void_p)))
        return out

    def mode_overlap(self, monitor_ids, direction: str, mode_fields) -> np.ndarray:
        """0.5 * sum((E_sim x H_mode*)_n + (E_mode x H_sim*)_n) per frequency over the plane the six DFT monitor ops
        `monitor_ids` (Ex..Hz order) share, reduced on the device (fdtd_mode_overlap).  mode_fields: six arrays (Ex..Hz)
        of the box shape."""
        op = self._mon_ops[monitor_ids[0]]
        m = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.complex128).reshape(op.shape) for a in mode_fields]))
        out = np.zeros(op.n_freq, dtype=np.complex128)
        ids = (C.c_int32 * 6)(*[int(i) for i in monitor_ids])
        _lib.check(self._lib.fdtd_mode_overlap(self._h, ids, "xyz".index(direction), m.ctypes.data_as(C.c_void_p),
                                               out.ctypes.data_as(C.c_void_p)))
        return out

    def set_dft(self, monitor_id: int, values) -> None:
        op = self._mon_ops[monitor_id]
        v = np.ascontiguousarray(values, dtype=np.complex128)
        if v.shape != (op.n_freq,) + op.shape:
            raise ValueError(f"DFT state shape {v.shape} != {(op.n_freq,) + op.shape}")
        if v.size:
            _lib.check(self._lib.fdtd_upload_dft(self._h, monitor_id, v.ctypes.data_as(C.c_void_p)))

    # ---- introspection -----------------------------------------------------------------------------------------------------
    @property
    def steps_done(self) -> int:
        v = C.c_int64()
        _lib.check(self._lib.fdtd_steps_done(self._h, C.byref(v)))
        return v.value

    @property
    def kernel_launches(self) -> int:
        v = C.c_int64()
        _lib.check(self._lib.fdtd_kernel_launches(self._h, C.byref(v)))
        return v.value

    def mem_info(self):
        f, t = C.c_int64(), C.c_int64()
        _lib.check(self._lib.fdtd_mem_info(self._h, C.byref(f), C.byref(t)))
        return f.value, t.value
