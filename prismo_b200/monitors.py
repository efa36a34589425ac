"""Monitor descriptions + result containers, interface-compatible with ``prismo.monitors``.

Same constructor arguments, public attributes, private result attributes and getters as
/root/reference/src/prismo/monitors/{base,field,dft,flux,mode_monitor}.py, so that
``prismo_b200.lowering`` can fill either these objects or the reference's own after a device run.
Sampling and DFT accumulation happen on the device; the getters below only post-process results.

Parity semantics: ``DFTMonitor``, ``FluxMonitor`` and ``ModeExpansionMonitor`` sample the
``[:10, :10]`` corner patch whatever their centre/size say, and only work in 2-D — exactly like the
reference.  ``region_correct=True`` (our extension, off by default) makes DFT/Flux monitors sample
their real region and lifts the 2-D restriction.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np

_ALL = ["Ex", "Ey", "Ez", "Hx", "Hy", "Hz"]


class Monitor:
    def __init__(self, center, size, name: Optional[str] = None):
        self.center, self.size = center, size
        self.name = name or f"{self.__class__.__name__}_{id(self)}"
        self._grid = None

    def initialize(self, grid) -> None:
        self._grid = grid


class FieldMonitor(Monitor):
    def __init__(self, center, size, components: Union[list, str] = "all", name=None, time_domain=True,
                 frequencies: Optional[list] = None):
        super().__init__(center, size, name)
        if components == "all":
            self.components = list(_ALL)
        elif components == "E":
            self.components = _ALL[:3]
        elif components == "H":
            self.components = _ALL[3:]
        else:
            for c in components:
                if c not in _ALL:
                    raise ValueError(f"Invalid field component: {c}")
            self.components = components
        self.time_domain = time_domain
        self.frequencies = frequencies if frequencies is not None else []
        self._time_data = {c: [] for c in self.components} if time_domain else {}
        self._time_points: list = []
        self._freq_data = {c: {f: None for f in self.frequencies} for c in self.components} if self.frequencies else {}
        self._component_shapes: dict = {}

    def initialize(self, grid) -> None:
        super().initialize(grid)
        from .grid import YeeGrid

        g = YeeGrid.like(grid)
        bounds = g.region_bounds(self.center, self.size)
        for c in self.components:
            self._component_shapes[c] = tuple(b - a for a, b in g.component_box(c, *bounds))
            for f in self.frequencies:
                self._freq_data[c][f] = np.zeros(self._component_shapes[c], dtype=np.complex128)

    def get_time_data(self, component):
        if not self.time_domain:
            raise ValueError("Time domain data not enabled for this monitor")
        if component not in self.components:
            raise ValueError(f"Component {component} not recorded by this monitor")
        return np.array(self._time_points), np.array(self._time_data[component])

    def get_frequency_data(self, component, frequency):
        if not self.frequencies:
            raise ValueError("Frequency domain data not enabled for this monitor")
        if component not in self.components:
            raise ValueError(f"Component {component} not recorded by this monitor")
        if frequency not in self.frequencies:
            raise ValueError(f"Frequency {frequency} not recorded by this monitor")
        return self._freq_data[component][frequency]

    _PAIRS = (("Ey", "Hz", 1), ("Ez", "Hy", -1), ("Ez", "Hx", 1), ("Ex", "Hz", -1), ("Ex", "Hy", 1), ("Ey", "Hx", -1))

    def get_power_flow(self, frequency=None):
        """Poynting sum over the recorded pairs (monitors/field.py:201-287)."""
        if frequency is None and not self.time_domain:
            raise ValueError("Time domain data not enabled for this monitor")
        if frequency is not None and frequency not in self.frequencies:
            raise ValueError(f"Frequency {frequency} not recorded by this monitor")
        if not (any(c[0] == "E" for c in self.components) and any(c[0] == "H" for c in self.components)):
            raise ValueError("Need both E and H components to calculate Poynting vector")
        s = 0
        for e, h, sign in self._PAIRS:
            if e in self.components and h in self.components:
                if frequency is None:
                    s += sign * self.get_time_data(e)[1] * self.get_time_data(h)[1]
                else:
                    s += 0.5 * sign * np.real(self.get_frequency_data(e, frequency)
                                              * np.conj(self.get_frequency_data(h, frequency)))
        return (np.array(self._time_points), s) if frequency is None else s


class DFTMonitor(Monitor):
    def __init__(self, center, size, frequencies, components=None, name=None, backend=None, region_correct=False):
        super().__init__(center, size, name)
        self.frequencies = np.array(frequencies)
        self.omega = 2 * np.pi * self.frequencies
        self.components = ["Ex", "Ey", "Ez"] if components is None else components
        self.region_correct = region_correct
        self._dft_data: dict = {}
        self._time_steps = 0
        self._dt = None

    def initialize(self, grid) -> None:
        super().initialize(grid)
        from .grid import YeeGrid

        g = YeeGrid.like(grid)
        bounds = g.region_bounds(self.center, self.size)
        for c in self.components:
            shape = (10, 10)                                    # placeholder shape, monitors/dft.py:106
            if self.region_correct:
                shape = tuple(b - a for a, b in g.component_box(c, *bounds))
            self._dft_data[c] = np.zeros((len(self.frequencies),) + shape, dtype=np.complex128)

    def get_frequency_data(self, component, frequency_index=None):
        if component not in self._dft_data:
            raise ValueError(f"Component {component} not monitored")
        if frequency_index is None:
            return {f: self._dft_data[component][i] for i, f in enumerate(self.frequencies)}
        return self._dft_data[component][frequency_index]

    def get_intensity(self, component, frequency_index):
        return np.abs(self.get_frequency_data(component, frequency_index)) ** 2

    def get_power_spectrum(self, component, normalize=True):
        spec = np.array([np.sum(self.get_intensity(component, i)) for i in range(len(self.frequencies))], dtype=float)
        if normalize and np.max(spec) > 0:
            spec = spec / np.max(spec)
        return spec

    def get_transmission_spectrum(self, reference_power=None):
        power = np.zeros(len(self.frequencies))
        for i in range(len(self.frequencies)):
            for c in ("Ex", "Ey", "Ez"):
                if c in self._dft_data:
                    power[i] += np.sum(self.get_intensity(c, i))
        if reference_power is not None:
            return power / reference_power
        return power / np.max(power) if np.max(power) > 0 else power


class FluxMonitor(Monitor):
    def __init__(self, center, size, direction, name=None, frequencies=None, backend=None, region_correct=False):
        super().__init__(center, size, name)
        self.direction = direction.lower()
        if self.direction not in ("x", "y", "z"):
            raise ValueError("direction must be 'x', 'y', or 'z'")
        self.frequencies = frequencies
        self.region_correct = region_correct
        self._power_flow_history: list = []
        self._time_history: list = []
        if frequencies is not None:
            self.omega = 2 * np.pi * np.array(frequencies)
            self._dft_ex = self._dft_ey = self._dft_ez = None
            self._dft_hx = self._dft_hy = self._dft_hz = None

    def initialize(self, grid) -> None:
        super().initialize(grid)
        if self.frequencies is not None:
            n = len(self.frequencies)
            for c in _ALL:
                setattr(self, "_dft_" + c.lower(), np.zeros((n, 10, 10), dtype=np.complex128))

    def _dA(self):
        dx, dy, dz = self._grid.spacing
        return {"x": dy * dz, "y": dx * dz, "z": dx * dy}[self.direction]

    def get_time_domain_power(self):
        return np.array(self._time_history), np.array(self._power_flow_history)

    def get_frequency_domain_power(self, frequency_index=None):
        """P(w) = 0.5 * sum Re(E x H*)_n dA over the sampled patch (monitors/flux.py:228-291)."""
        if self.frequencies is None:
            raise RuntimeError("No frequencies specified for this monitor")
        dA = self._dA()

        def one(i):
            ex, ey, ez = self._dft_ex[i], self._dft_ey[i], self._dft_ez[i]
            hx, hy, hz = self._dft_hx[i], self._dft_hy[i], self._dft_hz[i]
            s = {"x": 0.5 * np.real(ey * np.conj(hz) - ez * np.conj(hy)),
                 "y": 0.5 * np.real(ez * np.conj(hx) - ex * np.conj(hz)),
                 "z": 0.5 * np.real(ex * np.conj(hy) - ey * np.conj(hx))}[self.direction]
            return float(np.sum(s) * dA)

        if frequency_index is None:
            return np.array([one(i) for i in range(len(self.frequencies))])
        return one(frequency_index)

    def get_transmission(self, reference_power=None):
        if self.frequencies is not None:
            power = self.get_frequency_domain_power()
        else:
            p = self.get_time_domain_power()[1]
            power = np.mean(p) if len(p) > 0 else 0.0
        return power / reference_power if reference_power is not None else power


class ModeExpansionMonitor(Monitor):
    def __init__(self, center, size, modes, direction="x", frequencies=None, name=None, backend=None):
        super().__init__(center, size, name)
        self.modes, self.direction, self.frequencies = modes, direction.lower(), frequencies
        if self.direction not in ("x", "y", "z"):
            raise ValueError("direction must be 'x', 'y', or 'z'")
        self._mode_coeffs_time = {i: [] for i in range(len(modes))}
        self._time_points: list = []
        if frequencies is not None:
            self.omega = 2 * np.pi * np.array(frequencies)
            self._mode_coeffs_freq = {i: np.zeros(len(frequencies), dtype=complex) for i in range(len(modes))}

    def get_mode_coefficient(self, mode_index, domain="time"):
        if domain == "time":
            return np.array(self._mode_coeffs_time[mode_index])
        if self.frequencies is None:
            raise RuntimeError("No frequencies specified")
        return self._mode_coeffs_freq[mode_index]
