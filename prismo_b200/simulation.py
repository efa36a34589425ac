"""Host-side mirror of the reference's stepping interface, executing on the B200 engine.

Same names, constructor arguments, attributes, ordering and error behaviour as
  ElectromagneticFields   /root/reference/src/prismo/core/fields.py:20-144
  MaxwellUpdater          core/solver.py:24-165, :535-552
  FDTDSolver              core/solver.py:571-704
  Simulation              core/simulation.py:22-164
so tests written against the reference read the same here.  Field arrays are host NumPy mirrors
(user-visible, mutable in place); every ``step`` / ``run`` executes on the GPU through ``Session``.
There is no CPU stepping code in this module.
"""
from __future__ import annotations

import time as _time
from typing import Any, Optional, Union

import numpy as np

from .grid import COMPONENTS, GridSpec, YeeGrid
from .session import Session

EPS0 = 8.854187817e-12
MU0 = 4 * np.pi * 1e-7


class ElectromagneticFields:
    def __init__(self, grid: YeeGrid, dtype: Any = None, backend=None):
        self.grid = grid
        self.dtype = np.float64 if dtype is None else dtype
        self._fields = {c: np.zeros(grid.get_field_shape(c), dtype=self.dtype) for c in COMPONENTS}

    def __getitem__(self, component):
        if component not in self._fields:
            raise KeyError(f"Unknown field component: {component}")
        return self._fields[component]

    def __setitem__(self, component, value):
        if component not in self._fields:
            raise KeyError(f"Unknown field component: {component}")
        if np.isscalar(value):
            self._fields[component].fill(value)
            return
        value = np.asarray(value)
        if value.shape != self._fields[component].shape:
            raise ValueError(f"Shape mismatch for {component}: expected {self._fields[component].shape}, got {value.shape}")
        self._fields[component][:] = value

    Ex = property(lambda s: s._fields["Ex"])
    Ey = property(lambda s: s._fields["Ey"])
    Ez = property(lambda s: s._fields["Ez"])
    Hx = property(lambda s: s._fields["Hx"])
    Hy = property(lambda s: s._fields["Hy"])
    Hz = property(lambda s: s._fields["Hz"])

    def zero_fields(self):
        for a in self._fields.values():
            a.fill(0.0)

    def get_field_energy(self, region=None) -> float:
        """0.5*eps0*sum|E|^2 dV + 0.5*mu0*sum|H|^2 dV over whole arrays (core/fields.py:237-285)."""
        g = self.grid
        dv = g.dx * g.dy * (g.dz if g.is_3d else 1.0)
        e = sum(float(np.sum(self._fields[c] ** 2)) for c in ("Ex", "Ey", "Ez"))
        h = sum(float(np.sum(self._fields[c] ** 2)) for c in ("Hx", "Hy", "Hz"))
        return 0.5 * EPS0 * e * dv + 0.5 * MU0 * h * dv


class MaxwellUpdater:
    """Holds the cell-centred coefficients; ``step`` runs H pass + E pass on the device."""

    def __init__(self, grid: YeeGrid, dt: float, material_arrays: Optional[dict] = None, backend=None,
                 dtype=None, device=None, physics=None):
        self._physics = physics
        if backend is not None and not isinstance(backend, str) and not hasattr(backend, "zeros"):
            raise TypeError("backend must be a Backend instance or string name")
        self.grid, self.dt = grid, dt
        self.eps0, self.mu0, self.c = EPS0, MU0, 299792458.0
        courant = grid.get_courant_number(dt)
        if courant >= 1.0:
            raise ValueError(f"Time step dt={dt:.2e} violates Courant condition (S={courant:.3f} >= 1)")
        self.dx, self.dy, self.dz = grid.spacing
        self._dtype, self._device = dtype, device
        self._session: Optional[Session] = None
        self._set_materials(material_arrays)

    def _set_materials(self, m):
        dims = self.grid.dimensions
        if m is None:
            # vacuum: uniform scalars, no O(N) host arrays (the reference allocates eight, solver.py:84-133)
            self.uniform = True
            self.Ca, self.Cb, self.Da, self.Db = 1.0, self.dt / EPS0, 1.0, self.dt / MU0
            return
        self.uniform = False
        one, zero = np.ones(dims), np.zeros(dims)
        self.eps_rel = np.asarray(m.get("eps_rel", one))
        self.mu_rel = np.asarray(m.get("mu_rel", one))
        self.sigma_e = np.asarray(m.get("sigma_e", zero))
        self.sigma_m = np.asarray(m.get("sigma_m", zero))
        eps = EPS0 * self.eps_rel
        s = self.sigma_e * self.dt / (2 * eps)
        self.Ca = (1 - s) / (1 + s)
        self.Cb = (self.dt / eps) / (1 + s)
        mu = MU0 * self.mu_rel
        s = self.sigma_m * self.dt / (2 * mu)
        self.Da = (1 - s) / (1 + s)
        self.Db = (self.dt / mu) / (1 + s)

    def session(self) -> Session:
        if self._session is None:
            self._session = Session(self.grid, self.dt, dtype=self._dtype, device=self._device, physics=self._physics)
        self._session.set_coefficients(self.Ca, self.Cb, self.Da, self.Db)
        return self._session

    def update_magnetic_fields(self, fields) -> None:
        self.session().half_step(fields, "H")

    def update_electric_fields(self, fields) -> None:
        self.session().half_step(fields, "E")

    def step(self, fields) -> None:
        self.session().advance(fields, (), (), 0.0, self.dt, 1)

    def get_time_step(self):
        return self.dt

    def get_courant_number(self):
        return self.grid.get_courant_number(self.dt)


class FDTDSolver:
    def __init__(self, grid: YeeGrid, dt: Optional[float] = None, material_arrays: Optional[dict] = None,
                 backend=None, dtype=None, device=None, physics=None):
        self.grid = grid
        if dt is None:
            dt = grid.suggest_time_step(safety_factor=0.95)
        self._fields: Optional[ElectromagneticFields] = None      # allocated on first use (reference: eagerly, :611)
        self.updater = MaxwellUpdater(grid, dt, material_arrays, backend=backend, dtype=dtype, device=device,
                                      physics=physics)
        self.time, self.step_count = 0.0, 0

    @property
    def fields(self) -> ElectromagneticFields:
        if self._fields is None:
            self._fields = ElectromagneticFields(self.grid)
        return self._fields

    def _advance(self, fields, n, callback=None):
        dt = self.updater.get_time_step()
        if callback is None:
            self.updater.session().advance(fields, (), (), 0.0, dt, n)
            for _ in range(n):
                self.time += dt
            self.step_count += n
            return
        for step in range(n):                                       # callbacks see every step (solver.py:631-641)
            self.updater.session().advance(fields, (), (), 0.0, dt, 1)
            self.time += dt
            self.step_count += 1
            callback(self, step)

    def run(self, total_time: float, callback=None) -> None:
        self._advance(self.fields, int(np.ceil(total_time / self.updater.get_time_step())), callback)

    def run_steps(self, num_steps: int, callback=None) -> None:
        self._advance(self.fields, num_steps, callback)

    def step(self, fields: Optional[ElectromagneticFields] = None) -> None:
        self._advance(self.fields if fields is None else fields, 1)

    def reset(self) -> None:
        self.fields.zero_fields()
        self.time, self.step_count = 0.0, 0

    def get_simulation_info(self) -> dict:
        return {"time": self.time, "step_count": self.step_count, "dt": self.updater.get_time_step(),
                "courant_number": self.updater.get_courant_number(), "grid_dimensions": self.grid.dimensions,
                "is_2d": self.grid.is_2d, "field_energy": self.fields.get_field_energy()}


class Simulation:
    def __init__(self, size, resolution: Union[float, tuple], boundary_conditions: str = "pml", pml_layers: int = 10,
                 courant_factor: float = 0.9, dtype=None, device=None, physics=None):
        """``physics`` (our extension, default off): True or a ``PMLParams`` runs the stable Yee leap-frog with a
        working CPML instead of the reference's scheme — see DESIGN.md "physics mode"."""
        self.grid_spec = GridSpec(size=size, resolution=resolution, boundary_layers=pml_layers)
        self.grid = YeeGrid(self.grid_spec)
        self.size, self.resolution = size, resolution
        self.boundary_conditions, self.courant_factor = boundary_conditions, courant_factor
        self.fields = ElectromagneticFields(self.grid)
        self.dt = self.grid.get_time_step(courant_factor)
        self._dtype, self._device, self._physics = dtype, device, physics
        self.solver = FDTDSolver(self.grid, self.dt, dtype=dtype, device=device, physics=physics)
        self.sources: list = []
        self.monitors: list = []
        self._b200_ade: list = []
        self.step_count, self.current_time = 0, 0.0

    def add_ade(self, solver, component: str, mask=None) -> None:
        """Attach a dispersive medium's ADE recursion (materials.ADESolver), driven by E[component] each step."""
        from .materials import attach_ade

        attach_ade(self, solver, component, mask)

    def add_source(self, source) -> None:
        source.initialize(self.grid)
        self.sources.append(source)

    def add_monitor(self, monitor) -> None:
        monitor.initialize(self.grid)
        self.monitors.append(monitor)

    def set_materials(self, material_arrays: Optional[dict]) -> None:
        """Convenience for ``sim.solver = FDTDSolver(sim.grid, sim.dt, material_arrays)`` — the only way
        materials enter the reference."""
        self.solver = FDTDSolver(self.grid, self.dt, material_arrays, dtype=self._dtype, device=self._device,
                                 physics=self._physics)

    def set_geometry(self, shapes, background=None, coords=None) -> None:
        """Paint a shape list (geometry.Box / Sphere / Cylinder / Polygon / GeometryGroup, or the reference's own shape
        objects) into the update coefficients on the device instead of building material arrays on the host
        (geometry/shapes.py:71-99 + core/solver.py:113-133): Session.set_geometry."""
        self.solver.updater.session().set_geometry(shapes, background, coords)

    # ---- stepping -------------------------------------------------------------------------------------
    def run_steps(self, n: int) -> None:
        """n full steps in one device submission (solver, sources, monitors; core/simulation.py:147-164)."""
        if n <= 0:
            return
        sess = self.solver.updater.session()
        self.current_time = sess.advance(self.fields, self.sources, self.monitors, self.current_time, self.dt, n,
                                         ades=self._b200_ade)
        self.step_count += n
        dt_s = self.solver.updater.get_time_step()
        for _ in range(n):
            self.solver.time += dt_s
        self.solver.step_count += n

    def step(self) -> None:
        self.run_steps(1)

    def run(self, time: float, progress_callback=None, progress_interval: int = 10) -> None:
        steps = int(np.ceil(time / self.dt))
        start = _time.time()
        if progress_callback is None:
            self.run_steps(steps)
            return
        i = 0
        while i < steps:
            # the reference reports after step i whenever i % interval == 0 (core/simulation.py:136-139):
            # run up to and including the next such i on the device, then call back on the host
            nxt = i if i % progress_interval == 0 else (i // progress_interval + 1) * progress_interval
            nxt = min(nxt, steps - 1)
            self.run_steps(nxt - i + 1)
            if nxt % progress_interval == 0:
                progress_callback(nxt, steps, self.current_time, _time.time() - start)
            i = nxt + 1
        progress_callback(steps, steps, self.current_time, _time.time() - start)

    # ---- read-out helpers (core/simulation.py:166-250) -------------------------------------------------
    def get_field_data(self, monitor, component):
        if monitor not in self.monitors:
            raise ValueError("Monitor not found in this simulation")
        return monitor.get_time_data(component)[1]

    def get_frequency_data(self, monitor, component, frequency):
        if monitor not in self.monitors:
            raise ValueError("Monitor not found in this simulation")
        return monitor.get_frequency_data(component, frequency)

    def get_transmission(self, monitor, frequency=None):
        if monitor not in self.monitors:
            raise ValueError("Monitor not found in this simulation")
        if frequency is None:
            return np.mean(monitor.get_power_flow()[1])
        return np.mean(monitor.get_power_flow(frequency))
