"""Host post-processing of the tiny corner patches the placeholder monitors sample (<= 600 numbers/step).

The reference's Flux and ModeExpansion monitors reduce six 10x10 patches to one scalar per step
(monitors/flux.py:131-175; utils/mode_matching.py:41-171).  The device records the patches; these few
hundred flops per step are done here with the same NumPy expressions so the scalars match bit for bit.
"""
from __future__ import annotations

import numpy as np

_C = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


def patch_power(rec, step, direction, dA) -> float:
    """sum((E x H)_n) * dA on the recorded patches (monitors/flux.py:146-175)."""
    ex, ey, ez, hx, hy, hz = (rec[c][step] for c in _C)
    if direction == "x":
        s = ey * hz - ez * hy
    elif direction == "y":
        s = ez * hx - ex * hz
    else:
        s = ex * hy - ey * hx
    return float(np.sum(s) * dA)


def _resize(a, shape):
    if a.shape == shape:
        return a
    from scipy.ndimage import zoom

    zf = tuple(t / s for t, s in zip(shape, a.shape))
    if np.iscomplexobj(a):
        return zoom(a.real, zf, order=1) + 1j * zoom(a.imag, zf, order=1)
    return zoom(a, zf, order=1)


def _normal_poynting(e1, h2, e2, h1):
    return e1 * np.conj(h2) - e2 * np.conj(h1)


def mode_power(mode, direction="z", dx=1.0, dy=1.0) -> float:
    """|0.5 Re sum(E x H*)_n| dx dy (utils/mode_matching.py:134-171)."""
    d = direction.lower()
    if d == "x":
        s = _normal_poynting(mode.Ey, mode.Hz, mode.Ez, mode.Hy)
    elif d == "y":
        s = _normal_poynting(mode.Ez, mode.Hx, mode.Ex, mode.Hz)
    else:
        s = _normal_poynting(mode.Ex, mode.Hy, mode.Ey, mode.Hx)
    return float(abs(0.5 * np.real(np.sum(s)) * dx * dy))


def mode_overlap(six, mode, direction="z", dx=1.0, dy=1.0) -> complex:
    """0.5 * sum(S_sim + S_mode) dx dy / P_mode (utils/mode_matching.py:41-131)."""
    ex, ey, ez, hx, hy, hz = six
    m = {c: _resize(getattr(mode, c), ex.shape) for c in _C}
    d = direction.lower()
    if d == "x":
        s_sim = _normal_poynting(ey, m["Hz"], ez, m["Hy"])
        s_mode = _normal_poynting(m["Ey"], hz, m["Ez"], hy)
    elif d == "y":
        s_sim = _normal_poynting(ez, m["Hx"], ex, m["Hz"])
        s_mode = _normal_poynting(m["Ez"], hx, m["Ex"], hz)
    else:
        s_sim = _normal_poynting(ex, m["Hy"], ey, m["Hx"])
        s_mode = _normal_poynting(m["Ex"], hy, m["Ey"], hx)
    overlap = 0.5 * np.sum(s_sim + s_mode) * dx * dy
    power = mode_power(mode, direction, dx, dy)
    return complex(overlap / power) if abs(power) > 1e-20 else complex(0.0)


# ---- region-correct S-parameter extraction from DFT planes (extension) ----------------------------
def mode_coefficient_from_dft(flux_monitor, mode, frequency_index: int) -> complex:
    """Overlap of the frequency-domain fields a region-correct FluxMonitor accumulated on its plane with a waveguide
    mode (same formula as utils/mode_matching.py:41-131, evaluated on the monitor's real plane and with the grid's
    cell area instead of the placeholder's dx = dy = 1)."""
    six = []
    for c in _C:
        a = getattr(flux_monitor, "_dft_" + c.lower())[frequency_index]
        six.append(np.squeeze(a))
    d = flux_monitor.direction
    sp = flux_monitor._grid.spacing
    du, dv = [sp[a] for a in range(3) if "xyz"[a] != d][:2]
    return mode_overlap(tuple(six), mode, d, du, dv if dv else 1.0)


def s_parameter(coefficient_out: complex, coefficient_in: complex) -> complex:
    """S_ij = a_out / a_in for mode coefficients taken at the output and input port planes."""
    return coefficient_out / coefficient_in if coefficient_in != 0 else complex("nan")
