"""CPML parameter profiles for the opt-in physics mode (host side, a few 1-D vectors).

``PMLParams`` mirrors the reference dataclass (boundaries/pml.py:24-46) and the polynomial grading follows
boundaries/pml.py:117-151; the recursion coefficients are the textbook Roden–Gedney ones, INCLUDING the 1/eps0 the
reference leaves out (pml.py:196,222-233 — its b is ~1 and a ~0).  The reference never applies its CPML,
so nothing here has reference numbers to match.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

EPS0 = 8.854187817e-12
ETA0 = float(np.sqrt((4 * np.pi * 1e-7) / EPS0))


@dataclass
class PMLParams:
    thickness: int
    sigma_max: Optional[float] = None
    kappa_max: float = 15.0
    alpha_max: float = 0.0
    polynomial_order: int = 3


def axis_coefficients(n_cells: int, d: float, dt: float, p: PMLParams) -> np.ndarray:
    """(6, n_cells): b, a, 1/kappa at the E-update derivative positions (n + 1/2), then at the H-update ones (n).
    Identity (0, 0, 1) outside the two layers of ``p.thickness`` cells."""
    t, m = int(p.thickness), p.polynomial_order
    sigma_max = p.sigma_max if p.sigma_max is not None else 0.8 * (m + 1) / (ETA0 * d)
    out = np.zeros((6, n_cells))
    for half, row in ((0.5, 0), (0.0, 3)):
        x = np.arange(n_cells, dtype=np.float64) + half
        rho = np.where(x < t, (t - x) / t, np.where(x > n_cells - 1 - t, (x - (n_cells - 1 - t)) / t, 0.0))
        rho = np.clip(rho, 0.0, 1.0)
        inside = rho > 0
        sigma = sigma_max * rho ** m
        kappa = 1.0 + (p.kappa_max - 1.0) * rho ** m
        alpha = np.where(inside, p.alpha_max * (1.0 - rho) ** m, 0.0)
        b = np.exp(-(sigma / kappa + alpha) * dt / EPS0)
        denom = kappa * (sigma + kappa * alpha)
        a = np.where(inside & (denom > 0), sigma * (b - 1.0) / np.where(denom > 0, denom, 1.0), 0.0)
        out[row] = np.where(inside, b, 0.0)
        out[row + 1] = a
        out[row + 2] = np.where(inside, 1.0 / kappa, 1.0)
    return out


def coefficient_table(dims, spacing, dt, p: PMLParams) -> np.ndarray:
    """Concatenation over x, y, z of axis_coefficients — the layout fdtd_set_cpml expects."""
    return np.concatenate([axis_coefficients(n, d, dt, p).ravel() for n, d in zip(dims, spacing)])
