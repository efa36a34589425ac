"""Oracle restatement of the reference waveforms (sources/waveform.py:63-229).  Test infrastructure."""
from __future__ import annotations

import numpy as np


class CW:
    """A*sin(w t + phi) (waveform.py:96)."""

    def __init__(self, frequency, amplitude=1.0, phase=0.0):
        self.frequency, self.amplitude, self.phase = frequency, amplitude, phase
        self.omega = 2 * np.pi * frequency

    def __call__(self, t):
        return self.amplitude * np.sin(self.omega * t + self.phase)

    value = __call__


class GaussianPulse:
    """A*exp(-0.5*((t-t0)/tau)^2)*sin(w t + phi), t0 = 3 tau by default (waveform.py:131,148-152)."""

    def __init__(self, frequency, pulse_width, amplitude=1.0, phase=0.0, delay=None):
        self.frequency, self.pulse_width = frequency, pulse_width
        self.amplitude, self.phase = amplitude, phase
        self.omega = 2 * np.pi * frequency
        self.delay = delay if delay is not None else 3 * pulse_width

    def __call__(self, t):
        tau = (t - self.delay) / self.pulse_width
        env = np.exp(-0.5 * tau * tau)
        return self.amplitude * env * np.sin(self.omega * t + self.phase)

    value = __call__


class Ricker:
    """A*(1-2 tau^2)*exp(-tau^2), tau = pi f (t - t0), t0 = 1.5/f (waveform.py:191-193)."""

    def __init__(self, frequency, amplitude=1.0, delay=None):
        self.frequency, self.amplitude = frequency, amplitude
        self.delay = delay if delay is not None else 1.5 / frequency

    def __call__(self, t):
        tau = np.pi * self.frequency * (t - self.delay)
        t2 = tau * tau
        return self.amplitude * (1.0 - 2.0 * t2) * np.exp(-t2)

    value = __call__


class Custom:
    """amplitude * f(t), f a user callable; phase is the function's business (waveform.py:196-229)."""

    def __init__(self, waveform_func, amplitude=1.0):
        self.waveform_func, self.amplitude, self.phase = waveform_func, amplitude, 0.0

    def __call__(self, t):
        if isinstance(t, np.ndarray):
            return self.amplitude * np.array([self.waveform_func(ti) for ti in t])
        return self.amplitude * self.waveform_func(t)

    value = __call__


def make_waveform(frequency, pulse=True, pulse_width=None, amplitude=1.0, phase=0.0):
    """The pulse/CW switch used by every stock source constructor (e.g. plane_wave.py:73-86)."""
    if pulse:
        if pulse_width is None:
            raise ValueError("pulse_width must be provided for pulsed sources")
        return GaussianPulse(frequency, pulse_width, amplitude, phase)
    return CW(frequency, amplitude, phase)
