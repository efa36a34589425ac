"""Oracle restatement of the reference E/H curl updates.  Test infrastructure only.

Follows /root/reference/src/prismo/core/solver.py:
  coefficients            :113-133   (a1)
  coefficient averaging   :458-533   (a2)
  3-D H pass / E pass     :167-253 / :255-309   (a3, a4)
  2-D H pass / E pass     :311-397 / :399-456   (a5, a6)
  one step = H then E     :535-552   (a7)

The reference's scheme is bug-compatible on purpose: BOTH curls are forward
differences, the 2-D Hx update carries a sign error, the last planes/rows of the H arrays are
never written, and the 2-D H branches are gated on full-array reductions.  The operation ORDER
of every expression below is the reference's, so in fp64 this restatement is bit-identical to it
(pinned by tests/test_oracle_vs_reference.py and the committed goldens).

Fields are a dict name -> ndarray (C order, z contiguous), updated in place.
"""
from __future__ import annotations

import numpy as np

EPS0 = 8.854187817e-12
MU0 = 4 * np.pi * 1e-7


def coefficients(eps_rel, mu_rel, sigma_e, sigma_m, dt):
    """Cell-centred Ca, Cb, Da, Db (solver.py:113-133)."""
    eps = EPS0 * eps_rel
    se = sigma_e * dt / (2 * eps)
    Ca = (1 - se) / (1 + se)
    Cb = (dt / eps) / (1 + se)
    mu = MU0 * mu_rel
    sm = sigma_m * dt / (2 * mu)
    Da = (1 - sm) / (1 + sm)
    Db = (dt / mu) / (1 + sm)
    return Ca, Cb, Da, Db


def vacuum_coefficients(dims, dt):
    one = np.ones(dims, dtype=np.float64)
    zero = np.zeros(dims, dtype=np.float64)
    return coefficients(one, one, zero, zero, dt)


# ---- forward difference / neighbour mean helpers ------------------------------------------
def _sl(ndim, axis, s):
    out = [slice(None)] * ndim
    out[axis] = s
    return tuple(out)


def _fwd(a, axis, d):
    """(a[+1] - a) / d along axis; the division is a true division (solver.py:178 etc.)."""
    n = a.ndim
    return (a[_sl(n, axis, slice(1, None))] - a[_sl(n, axis, slice(None, -1))]) / d


def _mean2(a, axis):
    """0.5*(a + a[+1]) (solver.py:488-501, 503-533)."""
    n = a.ndim
    return 0.5 * (a[_sl(n, axis, slice(None, -1))] + a[_sl(n, axis, slice(1, None))])


def _mean4(a, ax0, ax1):
    """0.25*(a00 + a10 + a01 + a11), summed in the reference's order (solver.py:458-486, 513-518).

    ax0 < ax1; "a10" is the +1 neighbour along ax0.
    """
    n = a.ndim
    lo, hi = slice(None, -1), slice(1, None)

    def pick(s0, s1):
        idx = [slice(None)] * n
        idx[ax0], idx[ax1] = s0, s1
        return a[tuple(idx)]

    return 0.25 * (pick(lo, lo) + pick(hi, lo) + pick(lo, hi) + pick(hi, hi))


def _crop(a, shape):
    return a[tuple(slice(0, s) for s in shape)]


# ---- 3-D ------------------------------------------------------------------------------------
def update_h_3d(F, Da, Db, spacing):
    """solver.py:167-253.  Cells outside the common ranges are never touched."""
    dx, dy, dz = spacing
    Ex, Ey, Ez, Hx, Hy, Hz = (F[c] for c in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"))

    # Hx (nx-1, ny, nz): rows j < ny-1 and k < nz-1 (:191-205)
    cz_y = _fwd(Ez, 1, dy)          # (nx-1, ny-2, nz)
    cy_z = _fwd(Ey, 2, dz)          # (nx-1, ny,   nz-2)
    nyc = min(Hx.shape[1], cz_y.shape[1], cy_z.shape[1])
    nzc = min(Hx.shape[2], cz_y.shape[2], cy_z.shape[2])
    da = _mean2(Da, 0)[: Hx.shape[0], :nyc, :nzc]
    db = _mean2(Db, 0)[: Hx.shape[0], :nyc, :nzc]
    Hx[:, :nyc, :nzc] = da * Hx[:, :nyc, :nzc] - db * (cz_y[:, :nyc, :nzc] - cy_z[:, :nyc, :nzc])

    # Hy (nx, ny-1, nz) (:212-229)
    cx_z = _fwd(Ex, 2, dz)          # (nx,   ny-1, nz-2)
    cz_x = _fwd(Ez, 0, dx)          # (nx-2, ny-1, nz)
    nxc = min(Hy.shape[0], cx_z.shape[0], cz_x.shape[0])
    nzc = min(Hy.shape[2], cx_z.shape[2], cz_x.shape[2])
    da = _mean2(Da, 1)[:nxc, : Hy.shape[1], :nzc]
    db = _mean2(Db, 1)[:nxc, : Hy.shape[1], :nzc]
    Hy[:nxc, :, :nzc] = da * Hy[:nxc, :, :nzc] - db * (cx_z[:nxc, :, :nzc] - cz_x[:nxc, :, :nzc])

    # Hz (nx, ny, nz-1) (:236-253)
    cy_x = _fwd(Ey, 0, dx)          # (nx-2, ny,   nz-1)
    cx_y = _fwd(Ex, 1, dy)          # (nx,   ny-2, nz-1)
    nxc = min(Hz.shape[0], cy_x.shape[0], cx_y.shape[0])
    nyc = min(Hz.shape[1], cy_x.shape[1], cx_y.shape[1])
    da = _mean2(Da, 2)[:nxc, :nyc, : Hz.shape[2]]
    db = _mean2(Db, 2)[:nxc, :nyc, : Hz.shape[2]]
    Hz[:nxc, :nyc, :] = da * Hz[:nxc, :nyc, :] - db * (cy_x[:nxc, :nyc, :] - cx_y[:nxc, :nyc, :])


def update_e_3d(F, Ca, Cb, spacing):
    """solver.py:255-309.  Whole E arrays are written, from the H just produced.

    Cb may be a tuple (Cb_x, Cb_y, Cb_z): the OPT-IN per-component extension (diagonal permittivity tensor, the case
    materials/tensor.py:508-514 handles at function level).  The reference's solver has one Cb per cell; each
    component's array is averaged exactly as the scalar one is for that component.  PARITY UNPINNED for distinct
    arrays (no reference numbers exist); with three identical arrays it IS the reference's update."""
    dx, dy, dz = spacing
    Ex, Ey, Ez, Hx, Hy, Hz = (F[c] for c in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"))
    Cbx, Cby, Cbz = Cb if isinstance(Cb, (tuple, list)) else (Cb, Cb, Cb)

    ca = _crop(_mean4(Ca, 1, 2), Ex.shape)
    cb = _crop(_mean4(Cbx, 1, 2), Ex.shape)
    Ex[...] = ca * Ex + cb * (_fwd(Hz, 1, dy) - _fwd(Hy, 2, dz))

    ca = _crop(_mean4(Ca, 0, 2), Ey.shape)
    cb = _crop(_mean4(Cby, 0, 2), Ey.shape)
    Ey[...] = ca * Ey + cb * (_fwd(Hx, 2, dz) - _fwd(Hz, 0, dx))

    ca = _crop(_mean4(Ca, 0, 1), Ez.shape)
    cb = _crop(_mean4(Cbz, 0, 1), Ez.shape)
    Ez[...] = ca * Ez + cb * (_fwd(Hy, 0, dx) - _fwd(Hx, 1, dy))


# ---- 2-D ------------------------------------------------------------------------------------
def update_h_2d(F, Da, Db, spacing):
    """solver.py:311-397.  Da/Db are (nx, ny) here (the reference slices [:, :, 0])."""
    dx, dy, _ = spacing
    Ex, Ey, Ez, Hx, Hy, Hz = (F[c] for c in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"))

    # gate 1 (:321): NaN anywhere makes np.max NaN and the comparison False
    if np.max(np.abs(Ez)) > 0:
        nxh, nyh = Hx.shape
        nxe, nye = Ez.shape
        if nxe >= nxh and nye > 0:
            c = (Ez[:nxh, 1:] - Ez[:nxh, :-1]) / dy
            da = _mean2(Da, 0)
            db = _mean2(Db, 0)
            nyc = min(nyh, c.shape[1])
            # sign error kept (:338-341): "+ db * dEz/dy"
            Hx[:, :nyc] = da[:nxh, :nyc] * Hx[:, :nyc] + db[:nxh, :nyc] * c[:, :nyc]
        nxh, nyh = Hy.shape
        if nxe > 0 and nye >= nyh:
            c = (Ez[1:, :nyh] - Ez[:-1, :nyh]) / dx
            da = _mean2(Da, 1)
            db = _mean2(Db, 1)
            nxc = min(nxh, c.shape[0])
            Hy[:nxc, :] = da[:nxc, :nyh] * Hy[:nxc, :] - db[:nxc, :nyh] * c[:nxc, :]

    # gate 2 (:367)
    if np.max(np.abs(Ex)) > 0 or np.max(np.abs(Ey)) > 0:
        cey = np.zeros_like(Hz)
        if Ey.shape[0] > 1:
            n = min(Ey.shape[0] - 1, Hz.shape[0] - 1)
            cey[:n, :] = (Ey[1 : n + 1, :] - Ey[:n, :]) / dx
        cex = np.zeros_like(Hz)
        if Ex.shape[1] > 1:
            n = min(Ex.shape[1] - 1, Hz.shape[1] - 1)
            cex[:, :n] = (Ex[:, 1 : n + 1] - Ex[:, :n]) / dy
        nxh, nyh = Hz.shape
        # Hz is cell-centred in 2-D: coefficients are not averaged (:520-523)
        Hz[:, :] = Da[:nxh, :nyh] * Hz - Db[:nxh, :nyh] * (cey - cex)


def update_e_2d(F, Ca, Cb, spacing):
    """solver.py:399-456.  Unconditional."""
    dx, dy, _ = spacing
    Ex, Ey, Ez, Hx, Hy, Hz = (F[c] for c in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"))

    if Ez.shape[0] > 0 and Ez.shape[1] > 0:
        chy = (Hy[1:, :] - Hy[:-1, :]) / dx
        chx = (Hx[:, 1:] - Hx[:, :-1]) / dy
        n0, n1 = Ez.shape
        ca = _mean4(Ca, 0, 1)[:n0, :n1]
        cb = _mean4(Cb, 0, 1)[:n0, :n1]
        Ez[:, :] = ca * Ez + cb * (chy[:n0, :n1] - chx[:n0, :n1])

    if Hz.shape[0] > 0 and Hz.shape[1] > 0:
        if Ex.shape[0] > 0 and Ex.shape[1] > 0:
            c = (Hz[:, 1:] - Hz[:, :-1]) / dy
            n0, n1 = Ex.shape
            ca = _mean2(Ca, 1)[:n0, :n1]
            cb = _mean2(Cb, 1)[:n0, :n1]
            Ex[:, :] = ca * Ex + cb * c
        if Ey.shape[0] > 0 and Ey.shape[1] > 0:
            c = (Hz[1:, :] - Hz[:-1, :]) / dx
            n0, n1 = Ey.shape
            ca = _mean2(Ca, 0)[:n0, :n1]
            cb = _mean2(Cb, 0)[:n0, :n1]
            Ey[:, :] = ca * Ey - cb * c


# ---- drivers ----------------------------------------------------------------------------------
def update_h(F, Da, Db, spacing, is_2d):
    if is_2d:
        update_h_2d(F, Da[:, :, 0] if Da.ndim == 3 else Da, Db[:, :, 0] if Db.ndim == 3 else Db, spacing)
    else:
        update_h_3d(F, Da, Db, spacing)


def update_e(F, Ca, Cb, spacing, is_2d):
    if is_2d:
        if isinstance(Cb, (tuple, list)):
            raise ValueError("per-component Cb is a 3-D extension")
        update_e_2d(F, Ca[:, :, 0] if Ca.ndim == 3 else Ca, Cb[:, :, 0] if Cb.ndim == 3 else Cb, spacing)
    else:
        update_e_3d(F, Ca, Cb, spacing)


def step(F, coeffs, spacing, is_2d):
    """One MaxwellUpdater.step (solver.py:535-552): H pass, then E pass."""
    Ca, Cb, Da, Db = coeffs
    update_h(F, Da, Db, spacing, is_2d)
    update_e(F, Ca, Cb, spacing, is_2d)
