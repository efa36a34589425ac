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


def mode_coefficients(flux_monitor, mode) -> np.ndarray:
    """Mode coefficient at every frequency of a region-correct FluxMonitor.  While the monitor's DFT planes are still
    resident on the device (after the advance() that filled them, before the next one) the plane sums run there
    (fdtd_mode_overlap: 4 planes of the mode go up, n_freq complex numbers come down); otherwise on the host arrays."""
    d = flux_monitor.direction
    sp = flux_monitor._grid.spacing
    du, dv = [sp[a] for a in range(3) if "xyz"[a] != d][:2]
    dv = dv if dv else 1.0
    dev = getattr(flux_monitor, "_b200_device", None)
    nf = len(flux_monitor._dft_ex)
    if dev is not None and getattr(dev[0], "ops_epoch", -1) == dev[2] and getattr(dev[0], "_h", None):
        shape = np.squeeze(flux_monitor._dft_ex[0]).shape
        fields = [_resize(np.asarray(getattr(mode, c)), shape) for c in _C]
        box = flux_monitor._dft_ex[0].shape
        num = dev[0].mode_overlap(dev[1], d, [f.reshape(box) for f in fields])
        power = mode_power(mode, d, du, dv)
        return num * du * dv / power if abs(power) > 1e-20 else np.zeros(nf, dtype=np.complex128)
    return np.array([mode_coefficient_from_dft(flux_monitor, mode, i) for i in range(nf)])


def separate_forward_backward(coefficient_left, coefficient_right, neff, distance, wavelength):
    """Forward / backward amplitudes from two planes a known distance apart (utils/mode_matching.py:228-293):
    a_L = a_f + a_b,  a_R = a_f e^{i phi} + a_b e^{-i phi},  phi = 2 pi Re(neff) d / lambda."""
    phi = 2 * np.pi * np.real(neff) / wavelength * distance
    ep, em = np.exp(1j * phi), np.exp(-1j * phi)
    det = ep - em
    if abs(det) > 1e-10:
        return (coefficient_right - coefficient_left * em) / det, (coefficient_left * ep - coefficient_right) / det
    return coefficient_right, 0.0


def two_port_s_parameters(port1_pair, port2_pair, mode1, mode2, neff1, neff2, distance, wavelengths):
    """S11 and S21 per frequency from two pairs of region-correct FluxMonitors (left / right plane of each port), port 1
    excited: S11 = b1 / a1 and S21 = a2 / a1 (analysis/sparameters.py:88-122 with forward / backward mode amplitudes)."""
    cl1, cr1 = mode_coefficients(port1_pair[0], mode1), mode_coefficients(port1_pair[1], mode1)
    cl2, cr2 = mode_coefficients(port2_pair[0], mode2), mode_coefficients(port2_pair[1], mode2)
    s11, s21 = [], []
    for i, lam in enumerate(wavelengths):
        a1, b1 = separate_forward_backward(cl1[i], cr1[i], neff1, distance, lam)
        a2, _ = separate_forward_backward(cl2[i], cr2[i], neff2, distance, lam)
        s11.append(b1 / a1 if a1 != 0 else complex("nan"))
        s21.append(a2 / a1 if a1 != 0 else complex("nan"))
    return np.array(s11), np.array(s21)


def s_parameter(coefficient_out: complex, coefficient_in: complex) -> complex:
    """S_ij = a_out / a_in for mode coefficients taken at the output and input port planes."""
    return coefficient_out / coefficient_in if coefficient_in != 0 else complex("nan")


def fill_sparameter_analyzer(analyzer, excitation_port: int, amplitudes: dict) -> None:
    """Hand device-reduced mode amplitudes to the reference's own ``SParameterAnalyzer`` (analysis/sparameters.py:88-122),
    whose ``s_matrix`` then feeds its ``export_touchstone`` (:233) unchanged.  amplitudes: port -> (forward, backward)
    arrays over the analyzer's frequencies, as ``separate_forward_backward`` returns them.  ``add_mode_data`` computes
    backward / forward for the excited port and, for every other port, "forward" / "backward" of the dict it is given —
    documented there as transmitted / incident — so port i != j is passed its own forward amplitude over the excited
    port's forward (incident) amplitude."""
    inc = np.asarray(amplitudes[excitation_port][0], dtype=np.complex128)
    for port, (fwd, bwd) in amplitudes.items():
        if port == excitation_port:
            analyzer.add_mode_data(port, excitation_port, {"forward": inc, "backward": np.asarray(bwd, dtype=np.complex128)})
        else:
            analyzer.add_mode_data(port, excitation_port, {"forward": np.asarray(fwd, dtype=np.complex128), "backward": inc})
