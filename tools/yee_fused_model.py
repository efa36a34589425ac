"""Executable model of k_fused3d_yee (prismo_b200/csrc/fdtd_yee_fused.cuh): the kernel's tiling, register window,
row / lane exchanges, masks, CPML offsets, store predicates and psi ping-pong, statement by statement, with NumPy
arrays of shape (rows, lanes, V) standing in for the registers of one CTA.  Checked bitwise against oracle/yee.py
(tests/test_yee_fused_model.py) — it validates the ALGORITHM the CUDA kernel encodes (not its compilation), on the
CPU, before any GPU time is spent on it.  Warp width, owner counts and V are parameters so that small grids exercise
many tiles, segments and rims.
"""
from __future__ import annotations

import numpy as np

COMPS = ("ex", "ey", "ez", "hx", "hy", "hz")


def slab_index(n, N, t):
    n = np.asarray(n)
    return np.where(n < t, n, np.where(n >= N - t - 1, n - (N - 2 * t - 1), -1))


class Model:
    def __init__(self, dims, spacing, dt, coeffs, cpml_axes=None, thickness=0, TJ=3, W=8, own_lanes=6, V=2, lx=4):
        self.nx, self.ny, self.nz = dims
        self.sp, self.dt = spacing, dt
        self.ca, self.cb, self.da, self.db = coeffs
        self.t = thickness
        self.TJ, self.W, self.own, self.V, self.lx = TJ, W, own_lanes, V, lx
        self.pz = -(-self.nz // (W * V)) * (W * V)
        nx, ny, pz = self.nx, self.ny, self.pz
        self.F = [{c: np.zeros((nx + 4, ny, pz)) for c in COMPS} for _ in range(2)]
        self.cur = 0
        ns = 2 * thickness + 1
        self.cx = cpml_axes
        fam = (1, 2, 2, 0, 0, 1, 1, 2, 2, 0, 0, 1)
        shp = {0: (ns, ny, pz), 1: (nx + 1, ns, pz), 2: (nx + 1, ny, ns + 3)}
        self.psi = [[np.zeros(shp[f]) if thickness else None for f in fam] for _ in range(2)]

    # ---- host side ---------------------------------------------------------------------------------------------
    def upload(self, comp, a):
        A = self.F[self.cur][comp.lower()]
        A[...] = 0
        A[:a.shape[0], :a.shape[1], :a.shape[2]] = a

    def download(self, comp, shape):
        return self.F[self.cur][comp.lower()][:shape[0], :shape[1], :shape[2]].copy()

    def step(self):
        nx, ny = self.nx, self.ny
        vec_per_row = self.pz // self.V
        ntk = -(-vec_per_row // self.own)
        ntj = -(-ny // self.TJ)
        nseg = -(-nx // self.lx)
        src, dst = self.F[self.cur], self.F[self.cur ^ 1]
        pin, pout = self.psi[self.cur], self.psi[self.cur ^ 1]
        for seg in range(nseg):
            for tj in range(ntj):
                for tk in range(ntk):
                    self._cta(seg, tj, tk, src, dst, pin, pout)
        self.cur ^= 1

    # ---- one CTA -------------------------------------------------------------------------------------------------
    def _cta(self, seg, tj, tk, src, dst, pin, pout):
        nx, ny, nz, pz = self.nx, self.ny, self.nz, self.pz
        TJ, W, V = self.TJ, self.W, self.V
        R = TJ + 1
        dx, dy, dz = self.sp
        row = np.arange(R)[:, None, None]
        lane = np.arange(W)[None, :, None]
        el = np.arange(V)[None, None, :]
        j = tj * TJ + row + 0 * lane + 0 * el
        k0 = (tk * self.own + lane) * V + 0 * row
        ke = k0 + el
        i0 = seg * self.lx
        i1 = min(i0 + self.lx, nx)
        ld_ok = (j < ny) & (k0 < pz)
        owner = ld_ok & (row < TJ) & (lane < self.own)
        rim_row = (row == TJ) & (j == j)
        tpm = self.t
        syj = slab_index(j, ny, tpm) if tpm else np.full_like(j, -1)
        szk = slab_index(ke, nz, tpm) if tpm else np.full_like(ke, -1)

        def ldv(A, plane, ok, dj=0):
            """vector load of plane `plane`, row j+dj, elements k0..k0+V-1; zeros where not ok"""
            jj = np.clip(j + dj, 0, ny - 1)
            kk = np.clip(ke, 0, pz - 1)
            v = A[plane, jj, kk] if 0 <= plane < A.shape[0] else np.zeros(jj.shape)
            return np.where(ok, v, 0.0)

        def lds(A, plane, ok, dk):
            """scalar load of element k0+dk of row j (shape (R, W, 1))"""
            jj = np.clip(j[..., :1], 0, ny - 1)
            kk = np.clip(k0[..., :1] + dk, 0, pz - 1)
            return np.where(ok[..., :1], A[plane, jj, kk], 0.0)

        def cstep(d, q, cond, store, idx, axis, pos, n):
            """cpml_step on the elements where cond: psi <- b psi + a d ; returns ki d + psi (elsewhere d)"""
            b, a, ki = (self.cx[axis][pos + m] for m in range(3))
            n_ = np.clip(n, 0, len(b) - 1)
            idx = tuple(np.where(cond, x, 0) for x in idx)
            p_old = pin[q][idx]
            p = b[n_] * p_old + a[n_] * d
            w = cond & store
            pout[q][tuple(x[w] for x in idx)] = p[w]
            return np.where(cond, ki[n_] * d + p, d)

        zero = np.zeros((R, W, V))
        e0x, e0y, e0z, hpx, hpy, hpz = (zero.copy() for _ in range(6))
        if i0 > 0:
            e0y, e0z = ldv(src["ey"], i0 - 1, ld_ok), ldv(src["ez"], i0 - 1, ld_ok)
        e1x, e1y, e1z = (ldv(src[c], i0, ld_ok) for c in ("ex", "ey", "ez"))
        h1x, h1y, h1z = (ldv(src[c], i0, ld_ok) for c in ("hx", "hy", "hz"))

        for i in range(i0 - 1, i1):
            p = i + 1
            more = ld_ok & (i + 1 < i1)
            n_ex, n_ey, n_ez = (ldv(src[c], p + 1, more) for c in ("ex", "ey", "ez"))
            n_hx, n_hy, n_hz = (ldv(src[c], p + 1, more) for c in ("hx", "hy", "hz"))
            # shared-memory exchange: row r reads row r-1's (e1z, e1x) and row r+1's (hpz, hpx)
            ez_jm = np.concatenate([ldv(src["ez"], p, ld_ok & (j >= 1), dj=-1)[:1], e1z[:-1]], axis=0)
            ex_jm = np.concatenate([ldv(src["ex"], p, ld_ok & (j >= 1), dj=-1)[:1], e1x[:-1]], axis=0)
            hz_jp = np.concatenate([hpz[1:], zero[:1]], axis=0)
            hx_jp = np.concatenate([hpx[1:], zero[:1]], axis=0)
            # shuffles: k-1 from the previous lane's last element (lane 0: global), k+1 from the next lane's first
            ok0 = ld_ok & (k0 >= 1)
            ey_km = np.concatenate([lds(src["ey"], p, ok0, -1)[:, :1], e1y[:, :-1, V - 1:]], axis=1)
            ex_km = np.concatenate([lds(src["ex"], p, ok0, -1)[:, :1], e1x[:, :-1, V - 1:]], axis=1)
            hy_kp = np.concatenate([hpy[:, 1:, :1], hpy[:, -1:, :1]], axis=1)       # last lane: its own value (unused)
            hx_kp = np.concatenate([hpx[:, 1:, :1], hpx[:, -1:, :1]], axis=1)
            ey_k = np.concatenate([ey_km, e1y[..., :-1]], axis=2)                    # element e-1 (e = 0: previous lane)
            ex_k = np.concatenate([ex_km, e1x[..., :-1]], axis=2)
            hy_k = np.concatenate([hpy[..., 1:], hy_kp], axis=2)                     # element e+1 (last: next lane)
            hx_k = np.concatenate([hpx[..., 1:], hx_kp], axis=2)

            # ---- H+[p] ----
            sxp = int(slab_index(p, nx, tpm)) if tpm else -1
            st_h = owner & (p < i1)
            px1, pxm = p < nx - 1, 1 <= p <= nx - 2
            jm, jy1 = (j >= 1) & (j <= ny - 2), j < ny - 1
            km, kz1, kz0 = (ke >= 1) & (ke <= nz - 2), ke < nz - 1, ke < nz
            P, J, K = np.full_like(j, p), j, ke
            SX = np.full_like(j, sxp)
            hnx, hny, hnz = h1x.copy(), h1y.copy(), h1z.copy()
            c = px1 & jm & km
            d1 = cstep((e1z - ez_jm) / dy, 6, c & (syj >= 0), st_h, (P, syj, K), 1, 3, J) if tpm else (e1z - ez_jm) / dy
            d2 = cstep((e1y - ey_k) / dz, 7, c & (szk >= 0), st_h, (P, J, szk), 2, 3, K) if tpm else (e1y - ey_k) / dz
            hnx = np.where(c, self.da * h1x - self.db * (d1 - d2), hnx)
            c = pxm & jy1 & km
            d1 = cstep((e1x - ex_k) / dz, 8, c & (szk >= 0), st_h, (P, J, szk), 2, 3, K) if tpm else (e1x - ex_k) / dz
            d2 = cstep((e1z - e0z) / dx, 9, c & (SX >= 0), st_h, (SX, J, K), 0, 3, P) if tpm else (e1z - e0z) / dx
            hny = np.where(c, self.da * h1y - self.db * (d1 - d2), hny)
            c = pxm & jm & kz1
            d1 = cstep((e1y - e0y) / dx, 10, c & (SX >= 0), st_h, (SX, J, K), 0, 3, P) if tpm else (e1y - e0y) / dx
            d2 = cstep((e1x - ex_jm) / dy, 11, c & (syj >= 0), st_h, (P, syj, K), 1, 3, J) if tpm else (e1x - ex_jm) / dy
            hnz = np.where(c, self.da * h1z - self.db * (d1 - d2), hnz)
            for name, v in (("hx", hnx), ("hy", hny), ("hz", hnz)):
                dst[name][p, j[st_h], ke[st_h]] = v[st_h]

            # ---- E+[i] ----
            if i >= i0:
                sxq = int(slab_index(i, nx, tpm)) if tpm else -1
                qx1 = i < nx - 1
                st_e = owner
                Q = np.full_like(j, i)
                SQ = np.full_like(j, sxq)
                nr = ~rim_row
                nx_, ny_, nz_ = e0x.copy(), e0y.copy(), e0z.copy()
                c = jy1 & kz1 & nr
                d1 = cstep((hz_jp - hpz) / dy, 0, c & (syj >= 0), st_e, (Q, syj, K), 1, 0, J) if tpm else (hz_jp - hpz) / dy
                d2 = cstep((hy_k - hpy) / dz, 1, c & (szk >= 0), st_e, (Q, J, szk), 2, 0, K) if tpm else (hy_k - hpy) / dz
                nx_ = np.where(c, self.ca * e0x + self.cb * (d1 - d2), nx_)
                c = qx1 & kz1 & ld_ok & nr
                d1 = cstep((hx_k - hpx) / dz, 2, c & (szk >= 0), st_e, (Q, J, szk), 2, 0, K) if tpm else (hx_k - hpx) / dz
                d2 = cstep((hnz - hpz) / dx, 3, c & (SQ >= 0), st_e, (SQ, J, K), 0, 0, Q) if tpm else (hnz - hpz) / dx
                ny_ = np.where(c, self.ca * e0y + self.cb * (d1 - d2), ny_)
                c = qx1 & jy1 & kz0 & nr
                d1 = cstep((hny - hpy) / dx, 4, c & (SQ >= 0), st_e, (SQ, J, K), 0, 0, Q) if tpm else (hny - hpy) / dx
                d2 = cstep((hx_jp - hpx) / dy, 5, c & (syj >= 0), st_e, (Q, syj, K), 1, 0, J) if tpm else (hx_jp - hpx) / dy
                nz_ = np.where(c, self.ca * e0z + self.cb * (d1 - d2), nz_)
                for name, v in (("ex", nx_), ("ey", ny_), ("ez", nz_)):
                    dst[name][i, j[st_e], ke[st_e]] = v[st_e]

            # ---- rotate ----
            e0x, e0y, e0z = e1x, e1y, e1z
            e1x, e1y, e1z = n_ex, n_ey, n_ez
            hpx, hpy, hpz = hnx, hny, hnz
            h1x, h1y, h1z = n_hx, n_hy, n_hz
