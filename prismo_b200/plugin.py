"""``register()`` — make the B200 engine a backend of the real ``prismo`` package.

After ``prismo_b200.register()``:
    prismo.set_backend("b200", device_id=0)      # backends/backend_manager.py:139-186 gains one name
    sim = prismo.Simulation(...)                 # unchanged reference objects
    sim.add_source(...); sim.add_monitor(...)
    sim.run(t)                                   # core/simulation.py:107-164 executes on the GPU

The reference tree is never edited: its module attributes are re-bound at run time (the three re-exports
of ``set_backend`` / ``list_available_backends`` at backends/__init__.py:8 and prismo/__init__.py:29 too),
and the stepping methods are wrapped so that objects created under another backend keep the stock path.
"""
from __future__ import annotations

import time as _time

import numpy as np

from . import session as _session

_registered = False


def _lib_mod():
    from . import _lib

    return _lib


def _make_backend_class(Backend):
    class B200Backend(Backend):
        """Backend ABC implementation (backends/base.py:14-219).  Array factories hand out HOST NumPy
        arrays — the mirrors that user code and the reference's own set-up code read and write; stepping
        never goes through them (it runs in libfdtd_b200.so)."""

        def __init__(self, device_id: int = 0):
            from . import _lib

            _lib.load()                                   # fail loudly if the CUDA library is missing
            self.device_id = int(device_id)
            _session.configure(device=self.device_id)

        name = property(lambda self: "b200")
        is_gpu = property(lambda self: True)
        float32 = property(lambda self: np.float32)
        float64 = property(lambda self: np.float64)
        complex64 = property(lambda self: np.complex64)
        complex128 = property(lambda self: np.complex128)
        int32 = property(lambda self: np.int32)
        int64 = property(lambda self: np.int64)
        pi = property(lambda self: np.pi)

        # array factories: page-locked host memory (uploads / downloads are then plain DMA at PCIe line rate)
        def zeros(self, shape, dtype=None):
            a = _lib_mod().pinned_empty(shape, dtype or np.float64)
            a[...] = 0
            return a

        def ones(self, shape, dtype=None):
            a = _lib_mod().pinned_empty(shape, dtype or np.float64)
            a[...] = 1
            return a

        def empty(self, shape, dtype=None):
            return _lib_mod().pinned_empty(shape, dtype or np.float64)

        def array(self, data, dtype=None):
            return np.array(data, dtype=dtype)

        def asarray(self, data, dtype=None):
            return np.asarray(data, dtype=dtype)

        def to_numpy(self, array):
            return np.asarray(array)

        def copy(self, array):
            return np.copy(array)

        sqrt = staticmethod(np.sqrt)
        exp = staticmethod(np.exp)
        sin = staticmethod(np.sin)
        cos = staticmethod(np.cos)
        abs = staticmethod(np.abs)
        where = staticmethod(np.where)
        dot = staticmethod(np.dot)
        matmul = staticmethod(np.matmul)

        def sum(self, array, axis=None):
            return np.sum(array, axis=axis)

        def max(self, array, axis=None):
            return np.max(array, axis=axis)

        def min(self, array, axis=None):
            return np.min(array, axis=axis)

        def mean(self, array, axis=None):
            return np.mean(array, axis=axis)

        def fft(self, array, axis=-1):
            return np.fft.fft(array, axis=axis)

        def ifft(self, array, axis=-1):
            return np.fft.ifft(array, axis=axis)

        def fft2(self, array, axes=(-2, -1)):
            return np.fft.fft2(array, axes=axes)

        def ifft2(self, array, axes=(-2, -1)):
            return np.fft.ifft2(array, axes=axes)

        def synchronize(self):
            """cudaDeviceSynchronize on this backend's device (every engine stream included)."""
            m = _lib_mod()
            m.check(m.load().fdtd_device_sync(self.device_id))

        def get_memory_info(self):
            """cudaMemGetInfo of the device (the reference's CuPy backend reports the same two numbers)."""
            import ctypes as C

            m = _lib_mod()
            f, t = C.c_int64(), C.c_int64()
            m.check(m.load().fdtd_device_mem_info(self.device_id, C.byref(f), C.byref(t)))
            return {"backend": "b200", "device_id": self.device_id, "free_bytes": f.value, "total_bytes": t.value,
                    "used_bytes": t.value - f.value}

        def __repr__(self):
            return f"B200Backend(device_id={self.device_id})"

    return B200Backend


def _is_b200(obj) -> bool:
    b = getattr(obj, "backend", None)
    return getattr(b, "name", None) == "b200"


def _session_for(updater) -> "_session.Session":
    s = getattr(updater, "_b200_session", None)
    if s is None or s.dt != float(updater.dt):
        s = _session.Session(updater.grid, updater.dt)
        updater._b200_session = s
    s.set_coefficients(updater.Ca, updater.Cb, updater.Da, updater.Db)
    return s


def set_geometry(sim, shapes, background=None, coords=None) -> None:
    """Paint a list of the reference's own shape objects (prismo.geometry.shapes Box / Sphere / Cylinder / Polygon /
    GeometryGroup, material = shape.material) into the update coefficients ON THE DEVICE, for a reference ``Simulation``,
    ``FDTDSolver`` or ``MaxwellUpdater`` running on the "b200" backend.  Replaces the host pipeline
    rasterize (geometry/shapes.py:71-99) -> material arrays -> MaxwellUpdater(material_arrays=...) (core/solver.py:79-133);
    the updater's host Ca..Db stay as they are and are ignored until ``clear_geometry``."""
    upd = getattr(getattr(sim, "solver", sim), "updater", getattr(sim, "updater", sim))
    if not _is_b200(upd):
        raise RuntimeError('set_geometry needs the "b200" backend: prismo.set_backend("b200") before building the Simulation')
    _session_for(upd).set_geometry(shapes, background, coords)


def clear_geometry(sim) -> None:
    upd = getattr(getattr(sim, "solver", sim), "updater", getattr(sim, "updater", sim))
    _session_for(upd).clear_geometry()


def _advance_sim(sim, n: int) -> None:
    if n <= 0:
        return
    sess = _session_for(sim.solver.updater)
    sim.current_time = sess.advance(sim.fields, sim.sources, sim.monitors, sim.current_time, sim.dt, n,
                                    ades=getattr(sim, "_b200_ade", ()))
    sim.step_count += n
    dts = sim.solver.updater.get_time_step()
    for _ in range(n):
        sim.solver.time += dts
    sim.solver.step_count += n


def run_with_progress(advance, steps, callback, interval, now):
    """Device chunks between the reference's callback points (core/simulation.py:132-145)."""
    start = _time.time()
    if callback is None:
        advance(steps)
        return
    i = 0
    while i < steps:
        nxt = i if i % interval == 0 else (i // interval + 1) * interval
        nxt = min(nxt, steps - 1)
        advance(nxt - i + 1)
        if nxt % interval == 0:
            callback(nxt, steps, now(), _time.time() - start)
        i = nxt + 1
    callback(steps, steps, now(), _time.time() - start)


def register(prismo=None):
    """Install the ``"b200"`` backend into the (already importable) reference package.  Idempotent."""
    global _registered
    if prismo is None:
        import prismo  # noqa: F811  — the reference package must be importable by the caller
    from prismo.backends import backend_manager as bm
    from prismo.backends.base import Backend
    from prismo.core import simulation as sim_mod
    from prismo.core import solver as solver_mod

    if getattr(bm, "_b200_registered", False):
        return bm.B200Backend
    B200Backend = _make_backend_class(Backend)
    bm.B200Backend = B200Backend

    orig_set, orig_list = bm.set_backend, bm.list_available_backends

    def set_backend(backend: str, device_id: int = 0):
        if isinstance(backend, str) and backend.lower() == "b200":
            bm._CURRENT_BACKEND = B200Backend(device_id=device_id)
            return bm._CURRENT_BACKEND
        try:
            return orig_set(backend, device_id)
        except ValueError as e:
            if "Unknown backend" in str(e):
                raise ValueError(f"Unknown backend '{backend}'. Available backends: {list_available_backends()}") from None
            raise

    def list_available_backends():
        out = list(orig_list())
        try:
            from . import _lib

            _lib.load()
            out.append("b200")
        except OSError:
            pass
        return out

    set_backend.__doc__, list_available_backends.__doc__ = orig_set.__doc__, orig_list.__doc__
    for mod in (bm, prismo.backends, prismo):
        mod.set_backend = set_backend
        mod.list_available_backends = list_available_backends

    # ---- stepping entry points -------------------------------------------------------------------------
    Simulation, FDTDSolver, MaxwellUpdater = sim_mod.Simulation, solver_mod.FDTDSolver, solver_mod.MaxwellUpdater
    o_sim_step, o_sim_run = Simulation.step, Simulation.run
    o_sol_step, o_sol_run, o_sol_run_steps = FDTDSolver.step, FDTDSolver.run, FDTDSolver.run_steps
    o_upd_step, o_upd_h, o_upd_e = MaxwellUpdater.step, MaxwellUpdater.update_magnetic_fields, MaxwellUpdater.update_electric_fields

    def sim_step(self):
        if not _is_b200(self.solver.updater):
            return o_sim_step(self)
        _advance_sim(self, 1)

    def sim_run(self, time, progress_callback=None, progress_interval=10):
        if not _is_b200(self.solver.updater):
            return o_sim_run(self, time, progress_callback, progress_interval)
        steps = int(np.ceil(time / self.dt))
        run_with_progress(lambda n: _advance_sim(self, n), steps, progress_callback, progress_interval,
                          lambda: self.current_time)

    def _solver_advance(self, fields, n, callback):
        dt = self.updater.get_time_step()
        if callback is None:
            _session_for(self.updater).advance(fields, (), (), 0.0, dt, n)
            for _ in range(n):
                self.time += dt
            self.step_count += n
            return
        for step in range(n):
            _session_for(self.updater).advance(fields, (), (), 0.0, dt, 1)
            self.time += dt
            self.step_count += 1
            callback(self, step)

    def sol_step(self, fields=None):
        if not _is_b200(self.updater):
            return o_sol_step(self, fields)
        _solver_advance(self, self.fields if fields is None else fields, 1, None)

    def sol_run(self, total_time, callback=None):
        if not _is_b200(self.updater):
            return o_sol_run(self, total_time, callback)
        _solver_advance(self, self.fields, int(np.ceil(total_time / self.updater.get_time_step())), callback)

    def sol_run_steps(self, num_steps, callback=None):
        if not _is_b200(self.updater):
            return o_sol_run_steps(self, num_steps, callback)
        _solver_advance(self, self.fields, num_steps, callback)

    def upd_step(self, fields):
        if not _is_b200(self):
            return o_upd_step(self, fields)
        _session_for(self).advance(fields, (), (), 0.0, self.dt, 1)

    def upd_h(self, fields):
        if not _is_b200(self):
            return o_upd_h(self, fields)
        _session_for(self).half_step(fields, "H")

    def upd_e(self, fields):
        if not _is_b200(self):
            return o_upd_e(self, fields)
        _session_for(self).half_step(fields, "E")

    for cls, name, fn, orig in ((Simulation, "step", sim_step, o_sim_step), (Simulation, "run", sim_run, o_sim_run),
                                (FDTDSolver, "step", sol_step, o_sol_step), (FDTDSolver, "run", sol_run, o_sol_run),
                                (FDTDSolver, "run_steps", sol_run_steps, o_sol_run_steps),
                                (MaxwellUpdater, "step", upd_step, o_upd_step),
                                (MaxwellUpdater, "update_magnetic_fields", upd_h, o_upd_h),
                                (MaxwellUpdater, "update_electric_fields", upd_e, o_upd_e)):
        fn.__doc__, fn.__name__, fn.__qualname__ = orig.__doc__, orig.__name__, orig.__qualname__
        fn._b200_original = orig
        setattr(cls, name, fn)

    # ---- anisotropic tensor update (row a23, materials/tensor.py:482-588): function-level operator ------------------
    try:
        import importlib

        tensor_mod = importlib.import_module(prismo.__name__ + ".materials.tensor")
    except Exception:                                     # optional sub-package: nothing to patch
        tensor_mod = None
    if tensor_mod is not None:
        from .materials import tensor_update

        AU = tensor_mod.AnisotropicUpdater
        o_e, o_h = AU.update_e_from_curl_h, AU.update_h_from_curl_e

        def au_e(self, E, curl_H):
            if not _is_b200(self):
                return o_e(self, E, curl_H)
            m = self.material
            if m.is_diagonal:
                return tensor_update(E, curl_H, self.dt / self.eps0, (m.epsilon.xx, m.epsilon.yy, m.epsilon.zz),
                                     device=self.backend.device_id)
            return tensor_update(E, curl_H, self.dt / self.eps0, None, inverse=self.inv_epsilon, device=self.backend.device_id)

        def au_h(self, H, curl_E):
            if not _is_b200(self):
                return o_h(self, H, curl_E)
            m = self.material
            if m.is_diagonal:
                return tensor_update(H, curl_E, self.dt / self.mu0, (m.mu.xx, m.mu.yy, m.mu.zz), negative=True,
                                     device=self.backend.device_id)
            return tensor_update(H, curl_E, self.dt / self.mu0, None, inverse=self.inv_mu, negative=True,
                                 device=self.backend.device_id)

        for name, fn, orig in (("update_e_from_curl_h", au_e, o_e), ("update_h_from_curl_e", au_h, o_h)):
            fn.__doc__, fn.__name__, fn.__qualname__ = orig.__doc__, orig.__name__, orig.__qualname__
            fn._b200_original = orig
            setattr(AU, name, fn)

    bm._b200_registered = True
    _registered = True
    return B200Backend
