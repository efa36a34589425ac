"""Oracle restatement of the reference's ADE recursions.  Test infrastructure only.

Follows /root/reference/src/prismo/materials/dispersion.py:189-231 (Lorentz), :267-286 (Drude), :323-336 (Debye)
for the coefficients and materials/ade.py:116-160 for the recursions (evaluated left to right, both Lorentz
E terms use the SAME E, nothing feeds back into E — SURVEY F7, row a22).
"""
from __future__ import annotations

import numpy as np


def lorentz_coeffs(poles, dt):
    out = []
    for w0, de, g in poles:                      # (omega_0, delta_epsilon, gamma)
        denom = 4.0 + 2 * g * dt + w0 ** 2 * dt ** 2
        c0 = 2 * de * w0 ** 2 * dt ** 2 / denom
        out.append((c0, c0, (8.0 - 2 * w0 ** 2 * dt ** 2) / denom, -(4.0 - 2 * g * dt + w0 ** 2 * dt ** 2) / denom))
    return out


def drude_coeffs(omega_p, gamma, dt):
    e = np.exp(-gamma * dt)
    return omega_p ** 2 / gamma * (1.0 - e), e


def debye_coeffs(eps_inf, eps_s, tau, dt):
    e = np.exp(-dt / tau)
    return (eps_s - eps_inf) * (1.0 - e), e


class OAde:
    """kind: 'lorentz' (params = [(w0, de, g), ...]), 'drude' (omega_p, gamma), 'debye' (eps_inf, eps_s, tau)."""

    def __init__(self, kind, params, dt, shape, component, mask=None):
        self.kind, self.component, self.mask = kind, component, mask
        if kind == "lorentz":
            self.c = lorentz_coeffs(params, dt)
            self.P = [np.zeros(shape) for _ in self.c]
            self.Pp = [np.zeros(shape) for _ in self.c]
        elif kind == "drude":
            self.c = drude_coeffs(*params, dt)
            self.J = np.zeros(shape)
        else:
            self.c = debye_coeffs(*params, dt)
            self.P = np.zeros(shape)

    def update(self, F):
        E = F[self.component]
        if self.mask is not None:
            E = E * self.mask                                        # ade.py:298-302
        if self.kind == "lorentz":
            for i, (c0, c1, c2, c3) in enumerate(self.c):
                new = c0 * E + c1 * E + c2 * self.P[i] + c3 * self.Pp[i]      # ade.py:129-134
                self.Pp[i] = self.P[i].copy()
                self.P[i] = new
        elif self.kind == "drude":
            self.J = self.c[0] * E + self.c[1] * self.J                # ade.py:149
        else:
            self.P = self.c[0] * E + self.c[1] * self.P                # ade.py:160


# ---- coupled mode (OPT-IN extension of the engine, no reference counterpart: PARITY UNPINNED) -------------------------
# The reference never feeds the polarisation back into E.  The engine's opt-in "ade_coupled" mode does, with
# the reference's own recursions, applied at the BEGINNING of a step on the current E:
#     P+  = recursion(E^n, P, P_prev)                          (same left-to-right arithmetic as OAde.update)
#     J   = sum over the recursions driving a component of  (P+ - P) * (eps0/dt)   (Lorentz, Debye)   or   J+ * eps0
#           (Drude): the reference's recursion coefficients carry no eps0, so its P and J are in units of eps0
#     E^{n+1} = [reference E update](E^n, H^{n+1})  -  Cb_c * J        Cb_c: the 4-point mean the update itself uses
def coupled_step(F, coeffs, spacing, ades, dt, kernels):
    """One coupled step of the fields (no sources / monitors): mirrors ade_in_sweep + the E stage of fdtd_het.cuh /
    fdtd_fused.cuh operation for operation (fp64 results are compared bitwise)."""
    from .kernels import _crop, _mean4

    Ca, Cb, Da, Db = coeffs
    kj = 8.854187817e-12
    kp = kj / dt
    J = {}
    for a in ades:
        acc = J.setdefault(a.component, np.zeros(F[a.component].shape))
        if a.kind == "lorentz":
            old = [p.copy() for p in a.P]
            a.update(F)
            for o, n in zip(old, a.P):
                acc += (n - o) * kp
        elif a.kind == "drude":
            a.update(F)
            acc += a.J * kj
        else:
            old = a.P.copy()
            a.update(F)
            acc += (a.P - old) * kp
    kernels.update_h(F, Da, Db, spacing, False)
    kernels.update_e(F, Ca, Cb, spacing, False)
    axes = {"Ex": (1, 2), "Ey": (0, 2), "Ez": (0, 1)}
    for c, j in J.items():
        cb = _crop(_mean4(Cb, *axes[c]), F[c].shape)
        F[c] -= cb * j
