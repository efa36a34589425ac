"""Dispersive-material descriptions and ADE state, interface-compatible with the reference.

  LorentzPole / LorentzMaterial / DrudeMaterial / DebyeMaterial   materials/dispersion.py:120-336
  ADESolver (state container: P_current, P_previous, J_current)   materials/ade.py:16-160

Only parameters, the closed-form recursion coefficients and the state arrays live here; the recursion itself
runs on the device once the solver is attached to a simulation (``Simulation.add_ade`` or
``prismo_b200.attach_ade`` for a reference Simulation).  Like the reference, the recursion is driven by E but
does NOT feed back into the E update.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class LorentzPole:
    omega_0: float
    delta_epsilon: float
    gamma: float


class DispersiveMaterial:
    def __init__(self, epsilon_inf: float = 1.0, name: str = ""):
        self.epsilon_inf, self.name = epsilon_inf, name


class LorentzMaterial(DispersiveMaterial):
    def __init__(self, epsilon_inf, poles, name=""):
        super().__init__(epsilon_inf, name)
        self.poles = list(poles)

    def permittivity(self, omega):
        eps = self.epsilon_inf * np.ones_like(omega, dtype=complex)
        for p in self.poles:
            eps += p.delta_epsilon * p.omega_0 ** 2 / (p.omega_0 ** 2 - omega ** 2 - 1j * omega * p.gamma)
        return eps

    def get_ade_coefficients(self, dt):
        out = {"poles": [], "epsilon_inf": self.epsilon_inf}
        for p in self.poles:
            w0, g, de = p.omega_0, p.gamma, p.delta_epsilon
            denom = 4.0 + 2 * g * dt + w0 ** 2 * dt ** 2
            c0 = 2 * de * w0 ** 2 * dt ** 2 / denom
            out["poles"].append({"C0": c0, "C1": c0, "C2": (8.0 - 2 * w0 ** 2 * dt ** 2) / denom,
                                 "C3": -(4.0 - 2 * g * dt + w0 ** 2 * dt ** 2) / denom,
                                 "omega_0": w0, "gamma": g, "delta_epsilon": de})
        return out


class DrudeMaterial(DispersiveMaterial):
    def __init__(self, epsilon_inf, omega_p, gamma, name=""):
        super().__init__(epsilon_inf, name)
        self.omega_p, self.gamma = omega_p, gamma

    def permittivity(self, omega):
        return self.epsilon_inf - self.omega_p ** 2 / (omega ** 2 + 1j * omega * self.gamma)

    def get_ade_coefficients(self, dt):
        e = np.exp(-self.gamma * dt)
        return {"C0": self.omega_p ** 2 / self.gamma * (1.0 - e), "C1": e, "omega_p": self.omega_p,
                "gamma": self.gamma, "epsilon_inf": self.epsilon_inf}


class DebyeMaterial(DispersiveMaterial):
    def __init__(self, epsilon_inf, epsilon_s, tau, name=""):
        super().__init__(epsilon_inf, name)
        self.epsilon_s, self.tau = epsilon_s, tau

    def permittivity(self, omega):
        return self.epsilon_inf + (self.epsilon_s - self.epsilon_inf) / (1.0 + 1j * omega * self.tau)

    def get_ade_coefficients(self, dt):
        e = np.exp(-dt / self.tau)
        return {"C0": (self.epsilon_s - self.epsilon_inf) * (1.0 - e), "C1": e, "tau": self.tau,
                "epsilon_s": self.epsilon_s, "epsilon_inf": self.epsilon_inf}


class ADESolver:
    """State of one dispersive medium over a whole component array (same attributes as the reference's)."""

    def __init__(self, material, dt: float, grid_shape: tuple, backend=None):
        self.material, self.dt, self.grid_shape = material, dt, tuple(grid_shape)
        self.eps0 = 8.854187817e-12
        self.coeffs = material.get_ade_coefficients(dt)
        if "poles" in self.coeffs:
            self.P_current = [np.zeros(self.grid_shape) for _ in self.coeffs["poles"]]
            self.P_previous = [np.zeros(self.grid_shape) for _ in self.coeffs["poles"]]
        elif "omega_p" in self.coeffs:
            self.J_current = np.zeros(self.grid_shape)
        else:
            self.P_current = np.zeros(self.grid_shape)

    def get_polarization_current(self):
        if "poles" in self.coeffs:
            total = np.zeros(self.grid_shape)
            for p in self.P_current:
                total = total + p
            return total
        return self.J_current if "omega_p" in self.coeffs else self.P_current


def attach_ade(sim, solver, component: str, mask=None) -> None:
    """Run ``solver``'s recursion on the device after every step of ``sim``, driven by ``sim.fields[component]``
    (what calling ``solver.update_polarization(sim.fields[component] * mask)`` after each ``sim.step()`` does on the
    reference).  Works for our Simulation and, after ``register()``, for the reference's."""
    if component not in ("Ex", "Ey", "Ez"):
        raise ValueError("ADE recursions are driven by an E component")
    if not hasattr(sim, "_b200_ade"):
        sim._b200_ade = []
    sim._b200_ade.append((solver, component, mask))
