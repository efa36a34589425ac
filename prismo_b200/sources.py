"""Source descriptions, interface-compatible with the reference's ``prismo.sources`` classes.

These objects only DESCRIBE an injection (same constructor arguments and public attributes as
/root/reference/src/prismo/sources/{base,point,plane_wave,tfsf,gaussian,mode}.py); they hold no
``update_fields`` arithmetic.  ``prismo_b200.lowering`` turns them — or the reference's own objects,
which carry the same attributes — into device injection ops.
"""
from __future__ import annotations

from typing import Optional

from .waveform import Waveform, _stock


def _split_direction(direction: str):
    sign, axis = 1, direction
    if direction[:1] in "+-" and len(direction) > 1:
        sign, axis = (1 if direction[0] == "+" else -1), direction[1:]
    if axis.lower() not in ("x", "y", "z"):
        raise ValueError(f"Invalid direction: {direction}. Must be one of: x, y, z, +x, -x, +y, -y, +z, -z")
    return axis, sign


def _check_perpendicular(direction, polarization):
    if direction.lower() == polarization.lower():
        raise ValueError(f"Polarization ({polarization}) must be perpendicular to propagation direction ({direction})")


class Source:
    def __init__(self, center, size, name: Optional[str] = None, enabled: bool = True):
        self.center, self.size = center, size
        self.name = name or f"{self.__class__.__name__}_{id(self)}"
        self.enabled = enabled
        self._grid = None

    def initialize(self, grid) -> None:
        self._grid = grid

    def enable(self):
        self.enabled = True

    def disable(self):
        self.enabled = False


class PointSource(Source):
    def __init__(self, position, component, waveform: Waveform, name=None, enabled=True):
        super().__init__(center=position, size=(0, 0, 0), name=name, enabled=enabled)
        self.component, self.waveform = component, waveform


class ElectricDipole(PointSource):
    def __init__(self, position, polarization, frequency, pulse=True, pulse_width=None, amplitude=1.0,
                 phase=0.0, name=None, enabled=True):
        super().__init__(position, {"x": "Ex", "y": "Ey", "z": "Ez"}[polarization.lower()],
                         _stock(frequency, pulse, pulse_width, amplitude, phase), name, enabled)


class MagneticDipole(PointSource):
    def __init__(self, position, polarization, frequency, pulse=True, pulse_width=None, amplitude=1.0,
                 phase=0.0, name=None, enabled=True):
        super().__init__(position, {"x": "Hx", "y": "Hy", "z": "Hz"}[polarization.lower()],
                         _stock(frequency, pulse, pulse_width, amplitude, phase), name, enabled)


class PlaneWaveSource(Source):
    def __init__(self, center, size, direction, polarization, frequency, pulse=True, pulse_width=None,
                 amplitude=1.0, phase=0.0, name=None, enabled=True):
        super().__init__(center, size, name, enabled)
        self.direction, self.direction_sign = _split_direction(direction)
        _check_perpendicular(self.direction, polarization)
        self.polarization, self.frequency = polarization, frequency
        self.wavelength = 299792458.0 / frequency
        self.waveform = _stock(frequency, pulse, pulse_width, amplitude, phase)


class TFSFSource(Source):
    def __init__(self, center, size, direction, polarization, frequency, pulse=True, pulse_width=None,
                 amplitude=1.0, phase=0.0, angle=0.0, name=None, enabled=True):
        super().__init__(center, size, name, enabled)
        self.direction, self.direction_sign = _split_direction(direction)
        _check_perpendicular(self.direction, polarization)
        self.polarization, self.frequency, self.angle = polarization, frequency, angle
        self.wavelength = 299792458.0 / frequency
        self.waveform = _stock(frequency, pulse, pulse_width, amplitude, phase)


class GaussianBeamSource(Source):
    def __init__(self, center, size, direction, polarization, frequency, beam_waist, pulse=True,
                 pulse_width=None, amplitude=1.0, phase=0.0, name=None, enabled=True):
        super().__init__(center, size, name, enabled)
        _check_perpendicular(direction, polarization)
        self.direction, self.polarization = direction, polarization
        self.frequency, self.beam_waist = frequency, beam_waist
        self.wavelength = 299792458.0 / frequency
        self.waveform = _stock(frequency, pulse, pulse_width, amplitude, phase)


class ModeSource(Source):
    """``mode`` needs Ex..Hz (2-D complex), x, y, frequency; ``waveform`` needs ``.value(t)`` exactly as
    the reference demands (sources/mode.py:219; no stock reference waveform has it — SURVEY F9)."""

    def __init__(self, center, size, mode, direction, waveform, amplitude=1.0, phase=0.0, name=None):
        super().__init__(center, size, name)
        self.mode, self.direction, self.waveform = mode, direction, waveform
        self.amplitude, self.phase = amplitude, phase
        self.sign = +1 if direction[0] == "+" else -1
        self.axis = direction[-1].lower()
