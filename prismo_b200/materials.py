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


# ---- anisotropic (tensor) materials: row a23 ---------------------------------------------------------------------
class TensorComponents:
    """Relative tensor entries (scalars or arrays): xx, yy, zz, symmetric off-diagonals xy, xz, yz and optional
    yx, zx, zy for a non-symmetric tensor — ``materials/tensor.py:16-108``."""

    def __init__(self, xx, yy, zz, xy=0.0, xz=0.0, yz=0.0, yx=None, zx=None, zy=None):
        self.xx, self.yy, self.zz, self.xy, self.xz, self.yz = xx, yy, zz, xy, xz, yz
        self.yx, self.zx, self.zy = yx, zx, zy

    def is_diagonal(self) -> bool:
        return bool(all(np.all(c == 0) for c in (self.xy, self.xz, self.yz))
                    and all(c is None or np.all(c == 0) for c in (self.yx, self.zx, self.zy)))

    def is_symmetric(self) -> bool:
        if self.yx is None and self.zx is None and self.zy is None:
            return True
        pairs = ((self.xy, self.yx), (self.xz, self.zx), (self.yz, self.zy))
        return bool(all(np.allclose(a, a if b is None else b) for a, b in pairs))

    def to_full_matrix(self, backend=None) -> np.ndarray:
        yx = self.xy if self.yx is None else self.yx
        zx = self.xz if self.zx is None else self.zx
        zy = self.yz if self.zy is None else self.zy
        rows = ((self.xx, self.xy, self.xz), (yx, self.yy, self.yz), (zx, zy, self.zz))
        if np.isscalar(self.xx):
            return np.array(rows)
        out = np.zeros(np.asarray(self.xx).shape + (3, 3))
        for i, r in enumerate(rows):
            for j, c in enumerate(r):
                out[..., i, j] = c
        return out


class TensorMaterial:
    """``materials/tensor.py:110-318``: holds ε and μ tensors; inverses are formed on the host by NumPy."""

    def __init__(self, epsilon: TensorComponents, mu: TensorComponents = None, name: str = "", backend=None):
        self.epsilon = epsilon
        self.mu = TensorComponents(xx=1.0, yy=1.0, zz=1.0) if mu is None else mu
        self.name = name or "AnisotropicMaterial"
        self.epsilon_tensor = epsilon.to_full_matrix()
        self.mu_tensor = self.mu.to_full_matrix()
        self.is_diagonal = epsilon.is_diagonal() and self.mu.is_diagonal()
        self.is_symmetric = epsilon.is_symmetric() and self.mu.is_symmetric()

    @staticmethod
    def _inverse(t: np.ndarray, comps: TensorComponents, diagonal: bool) -> np.ndarray:
        if diagonal:
            inv = np.zeros_like(t)
            inv[..., 0, 0], inv[..., 1, 1], inv[..., 2, 2] = 1.0 / comps.xx, 1.0 / comps.yy, 1.0 / comps.zz
            return inv
        return np.linalg.inv(t)

    def get_inverse_epsilon(self) -> np.ndarray:
        return self._inverse(self.epsilon_tensor, self.epsilon, self.is_diagonal)

    def get_inverse_mu(self) -> np.ndarray:
        return self._inverse(self.mu_tensor, self.mu, self.is_diagonal)


def _strong_dtype(x):
    """dtype a value contributes under NumPy's promotion rules (NEP 50): Python scalars are weak (None)."""
    if isinstance(x, (np.ndarray, np.generic)):
        return x.dtype
    return None


def _promote(*dts):
    """result dtype of float arithmetic between the given strong dtypes (None = weak Python scalar)."""
    strong = [d for d in dts if d is not None]
    rt = np.result_type(*strong) if strong else np.dtype(np.float64)
    return rt if rt.kind == "f" else np.result_type(rt, np.float64)


def tensor_update(fields, curls, scale: float, diagonal, inverse=None, negative: bool = False, device: int = None):
    """Device evaluation of the reference's tensor update on co-located host arrays (``materials/tensor.py:482-588``).

    diagonal = (d_x, d_y, d_z) scalars or arrays → ``f (+|-) ((scale * curl) / d)``; or ``inverse`` = (..., 3, 3) inverse
    tensor → ``f + (±scale) * ((r0*c0 + r1*c1) + r2*c2)``.  Result dtypes AND the precision of every intermediate
    follow NumPy promotion exactly as the reference's expressions do: Python-float entries are weak, so a float32
    curl is multiplied (and divided) in float32 even when the field or a tensor array then widens the sum to float64.
    Returns three new arrays."""
    import ctypes as C

    from . import _lib
    from .session import _state

    lib = _lib.load()
    f = [np.asarray(a) for a in fields]
    c = [np.asarray(a) for a in curls]
    shape = np.broadcast_shapes(*(a.shape for a in f + c))
    full = inverse is not None
    f32, f64 = np.dtype(np.float32), np.dtype(np.float64)
    if full:
        inv = np.asarray(inverse)
        entries = [inv[..., i, j] for i in range(3) for j in range(3)]
        rt = _promote(inv.dtype, *(a.dtype for a in f + c))
        if rt != f64:
            raise TypeError(f"tensor_update: full-tensor update promotes to {rt}; only float64 is supported")
        groups = {(f64, 0): (0, 1, 2)}
    else:
        entries = list(diagonal)
        groups = {}
        for k in range(3):
            mul = _promote(c[k].dtype)                                   # scale is a weak Python float
            div = _promote(mul, _strong_dtype(entries[k]))
            res = _promote(f[k].dtype, div)
            if res not in (f32, f64):
                raise TypeError(f"tensor_update: unsupported result dtype {res}")
            mode = (2 if (res == f64 and mul == f32) else 0) | (4 if (res == f64 and div == f32) else 0)
            groups.setdefault((res, mode), []).append(k)
    out = [None] * 3
    n = int(np.prod(shape, dtype=np.int64))
    for (dt, mode), comps in groups.items():

        def flat(a):
            return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=dt), shape)).reshape(-1)

        keep, fp, cp, op = [], (C.c_void_p * 3)(), (C.c_void_p * 3)(), (C.c_void_p * 3)()
        for k in range(3):
            if k in comps:
                a = flat(f[k]); o = np.empty(n, dtype=dt)
                keep += [a, o]; fp[k] = a.ctypes.data; op[k] = o.ctypes.data; out[k] = o.reshape(shape)
            if k in comps or full:
                b = flat(c[k]); keep.append(b); cp[k] = b.ctypes.data
        coef = (C.c_double * 9)()
        arrs = (C.c_void_p * 9)()
        for q, e in enumerate(entries):
            if np.ndim(e) == 0:
                coef[q] = float(e)
            else:
                a = flat(e); keep.append(a); arrs[q] = a.ctypes.data
        s = -scale if (full and negative) else scale
        _lib.check(lib.fdtd_tensor_update(_state["device"] if device is None else device,
                                          _lib.F32 if dt == f32 else _lib.F64, n, fp, cp, op, float(s),
                                          int(bool(negative) and not full), mode | (1 if full else 0), coef, arrs))
    return tuple(out)


class AnisotropicUpdater:
    """``materials/tensor.py:438-588`` on the device: same constructor, attributes and method names; accepts the
    reference's ``TensorMaterial`` as well as ours (duck-typed: ``is_diagonal``, ``epsilon``/``mu`` components,
    ``get_inverse_epsilon``/``get_inverse_mu``)."""

    def __init__(self, tensor_material, dt: float, backend=None, device: int = None):
        if backend is not None and not isinstance(backend, str) and not hasattr(backend, "zeros"):
            raise TypeError("backend must be a Backend instance or string name")
        self.material, self.dt, self._device = tensor_material, dt, device
        self.eps0, self.mu0 = 8.854187817e-12, 4 * np.pi * 1e-7
        self.inv_epsilon = tensor_material.get_inverse_epsilon()
        self.inv_mu = tensor_material.get_inverse_mu()

    def update_e_from_curl_h(self, E, curl_H):
        m = self.material
        if m.is_diagonal:
            return tensor_update(E, curl_H, self.dt / self.eps0, (m.epsilon.xx, m.epsilon.yy, m.epsilon.zz), device=self._device)
        return tensor_update(E, curl_H, self.dt / self.eps0, None, inverse=self.inv_epsilon, device=self._device)

    def update_h_from_curl_e(self, H, curl_E):
        m = self.material
        if m.is_diagonal:
            return tensor_update(H, curl_E, self.dt / self.mu0, (m.mu.xx, m.mu.yy, m.mu.zz), negative=True, device=self._device)
        return tensor_update(H, curl_E, self.dt / self.mu0, None, inverse=self.inv_mu, negative=True, device=self._device)
