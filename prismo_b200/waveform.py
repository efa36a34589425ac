"""Temporal waveforms, interface-compatible with the reference (sources/waveform.py:63-229).

The engine never evaluates these on the device: the host tabulates ``waveform(t_n)`` once per run and
uploads the table (see lowering.py), so user-defined callables work unchanged.
"""
from __future__ import annotations

import numpy as np


class Waveform:
    def __init__(self, amplitude=1.0, phase=0.0):
        self.amplitude, self.phase = amplitude, phase

    def __call__(self, t):
        return self.evaluate(t)

    def evaluate(self, t):
        raise NotImplementedError("Subclasses must implement evaluate method")


class ContinuousWave(Waveform):
    def __init__(self, frequency, amplitude=1.0, phase=0.0):
        super().__init__(amplitude, phase)
        self.frequency = frequency
        self.omega = 2 * np.pi * frequency

    def evaluate(self, t):
        return self.amplitude * np.sin(self.omega * t + self.phase)


class GaussianPulse(Waveform):
    def __init__(self, frequency, pulse_width, amplitude=1.0, phase=0.0, delay=None):
        super().__init__(amplitude, phase)
        self.frequency, self.pulse_width = frequency, pulse_width
        self.omega = 2 * np.pi * frequency
        self.delay = delay if delay is not None else 3 * pulse_width

    def evaluate(self, t):
        tau = (t - self.delay) / self.pulse_width
        return self.amplitude * np.exp(-0.5 * tau * tau) * np.sin(self.omega * t + self.phase)


class RickerWavelet(Waveform):
    def __init__(self, frequency, amplitude=1.0, delay=None):
        super().__init__(amplitude, 0.0)
        self.frequency = frequency
        self.delay = delay if delay is not None else 1.5 / frequency

    def evaluate(self, t):
        tau = np.pi * self.frequency * (t - self.delay)
        tau2 = tau * tau
        return self.amplitude * (1.0 - 2.0 * tau2) * np.exp(-tau2)


class CustomWaveform(Waveform):
    def __init__(self, waveform_func, amplitude=1.0):
        super().__init__(amplitude, 0.0)
        self.waveform_func = waveform_func

    def evaluate(self, t):
        if isinstance(t, np.ndarray):
            return self.amplitude * np.array([self.waveform_func(ti) for ti in t])
        return self.amplitude * self.waveform_func(t)


def _stock(frequency, pulse, pulse_width, amplitude, phase):
    if pulse:
        if pulse_width is None:
            raise ValueError("pulse_width must be provided for pulsed sources")
        return GaussianPulse(frequency=frequency, pulse_width=pulse_width, amplitude=amplitude, phase=phase)
    return ContinuousWave(frequency=frequency, amplitude=amplitude, phase=phase)
