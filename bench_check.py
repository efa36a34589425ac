"""Self-check of a bench run (the CHECKER side of bench.py, not part of the timed path).

After the timed region bench.py re-seeds the fields, runs a few steps of the same workload and calls into this module:

  * crop_rel_l2   the device fields on two cropped sub-domains (one containing the TFSF source plane, one starting at the
                  DFT monitor plane) against the oracle (oracle/kernels.py, the NumPy restatement of the reference update)
                  advanced from the same initial crop.  Locality makes this exact: both curls of the reference scheme are
                  forward differences, so after n steps a cell depends only on the cells at +0 .. +2n along every axis.
  * dft_rel_l2    the monitor's running DFT over the crop of its plane against  sum_s (F_s * phasor_s) * dt  of the
                  oracle crop (monitors/field.py:124-143 arithmetic).
  * fields_sha    sha256 over order- and decomposition-independent per-plane checksums of all six arrays
                  (fdtd_field_checksum): identical at 1/2/4/8 GPUs iff the fields are bitwise identical.

The initial fields are a pure function of the GLOBAL cell index (seed_plane x x_profile), so every rank of every
decomposition seeds the same global state.  Only bench.py and tests import this file; it imports oracle/ as the checker.
"""
from __future__ import annotations

import hashlib

import numpy as np

COMPONENTS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
SHORT = {"Ex": (1, 2), "Ey": (0, 2), "Ez": (0, 1), "Hx": (0,), "Hy": (1,), "Hz": (2,)}
N_CHECK = 4


def comp_shape(c, dims):
    s = list(dims)
    for ax in SHORT[c]:
        s[ax] -= 1
    return tuple(s)


def seed_planes(dims, scale=1e-3, seed=0):
    """One white-noise (j,k) plane per component (float32), drawn in component order from one generator."""
    rng = np.random.default_rng(seed)
    out = {}
    for c in COMPONENTS:
        shp = comp_shape(c, dims)
        out[c] = (rng.standard_normal(shp[1:]) * scale * (1.0 if c[0] == "E" else 1 / 377.0)).astype(np.float32)
    return out


def x_profile(c, i_global):
    """Modulation along x as a function of the GLOBAL plane index (float32): makes d/dx terms non-trivial."""
    i = np.asarray(i_global, dtype=np.float64)
    return (1.0 + 0.5 * np.cos(0.37 * i + COMPONENTS.index(c))).astype(np.float32)


def seed_box(planes, c, lo, hi):
    """Initial values of component c on the global box [lo, hi): float32 product, exactly what seed_fields uploads."""
    a = x_profile(c, np.arange(lo[0], hi[0]))
    p = planes[c][lo[1]:hi[1], lo[2]:hi[2]]
    return a[:, None, None] * p[None, :, :]


def seed_fields(eng, dims, x0=0, planes=None):
    """Upload the seeded state of the local slab [x0, x0 + local planes) (float32 values in either engine dtype)."""
    planes = planes or seed_planes(dims)
    for c in COMPONENTS:
        shp = eng.field_shape(c)
        a = x_profile(c, np.arange(x0, x0 + shp[0]))
        buf = np.empty(shp, dtype=np.float32)
        np.multiply(a[:, None, None], planes[c][None, :, :], out=buf)
        eng.upload(c, buf)
    return planes


def crop_plan(dims, src_plane, mon_plane, n=N_CHECK):
    """Two global crops [lo, lo + size) whose oracle blocks [lo, lo + size + 2n + 2) stay clear of the high faces."""
    nx, ny, nz = dims
    halo = 2 * n + 2
    size = (min(12, nx // 8), min(16, ny // 6), min(16, nz // 6))
    crops = []
    a_lo = (max(src_plane - 3, 0), ny // 2 - size[1] // 2, nz // 2 - size[2] // 2)
    b_lo = (mon_plane, ny // 3, nz // 3)
    for lo in (a_lo, b_lo):
        if all(l >= 0 and l + s + halo <= d - 1 for l, s, d in zip(lo, size, dims)):
            crops.append((tuple(int(v) for v in lo), size))
    return crops


def oracle_crop(planes, dims, dt, spacing, lo, size, amp, phasors, src_plane, mon_plane, n=N_CHECK, coef_fn=None,
                src_profile=None):
    """Advance the crop with the oracle; returns the final crop fields and, if the crop starts at the monitor plane, the
    DFT sums of (Ey, Hz) over its first plane.  coef_fn(lo, sdims) -> (Ca, Cb, Da, Db) of the crop for heterogeneous
    media (None: vacuum); src_profile: {component: global (ny', nz') profile} of a profiled plane source (None: the
    uniform TFSF plane)."""
    from oracle import kernels

    sdims = tuple(s + 2 * n + 2 for s in size)
    sub = {}
    for c in COMPONENTS:
        shp = comp_shape(c, sdims)
        sub[c] = seed_box(planes, c, lo, tuple(l + s for l, s in zip(lo, shp))).astype(np.float64)
    coeffs = kernels.vacuum_coefficients(sdims, dt) if coef_fn is None else coef_fn(lo, sdims)
    isrc = src_plane - lo[0]
    dft = None
    if lo[0] == mon_plane:
        dft = {c: np.zeros((phasors.shape[1],) + size[1:], dtype=np.complex128) for c in ("Ey", "Hz")}
    for s in range(n):
        kernels.step(sub, coeffs, spacing, False)
        if 0 <= isrc < sdims[0] - 1:                       # TFSF plane (tfsf.py:333-338): whole-plane injection
            for col, c in enumerate(("Ey", "Hz")):
                if src_profile is None:
                    sub[c][isrc] += amp[s, col]
                else:                                      # mode-source style: amplitude x transverse profile (mode.py:255-361)
                    shp = sub[c].shape
                    sub[c][isrc] += amp[s, col] * src_profile[c][lo[1]:lo[1] + shp[1], lo[2]:lo[2] + shp[2]]
        if dft is not None:
            for c in dft:
                d = sub[c][0, :size[1], :size[2]]
                for f in range(phasors.shape[1]):
                    dft[c][f] += d * phasors[s, f] * dt
    want = {c: sub[c][:size[0], :size[1], :size[2]].copy() for c in COMPONENTS}
    return want, dft


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / den) if den > 0 else float(np.linalg.norm(a.ravel()))


def sha_of_checksums(per_comp):
    """per_comp: {component: (global planes, 2) uint64} -> hex digest."""
    h = hashlib.sha256()
    for c in COMPONENTS:
        h.update(np.ascontiguousarray(per_comp[c], dtype=np.uint64).tobytes())
    return h.hexdigest()[:32]


def run_check(dims, dt, spacing, dtype, tables_fn, src_plane, mon_plane, reseed, run_steps, fetch_box, fetch_dft,
              fetch_checksums, do_oracle=True, n=N_CHECK, coef_fn=None, src_profile=None):
    """Drive the check through callables so the single-GPU and the slab-decomposed bench share it.

      reseed()                 upload the seeded state, zero the DFT sums, install tables for n steps from t = 0
      run_steps(n)             advance n steps (all ranks)
      fetch_box(c, lo, hi)     global box of component c as fp64 (rank 0 gets the assembled array, others None)
      fetch_dft(c, lo2, hi2)   (n_freq, j, k) crop of the monitor's DFT of component c (rank 0), or None
      fetch_checksums()        {component: (global planes, 2) uint64} on rank 0, None elsewhere
    """
    amp, ph, _ = tables_fn(n, dt)
    planes = reseed(n, amp, ph)
    run_steps(n)
    out = {"steps": n, "crop_rel_l2": None, "dft_rel_l2": None, "fields_sha": None, "crops": []}
    worst, worst_dft, exact = 0.0, None, True
    if do_oracle:
        for lo, size in crop_plan(dims, src_plane, mon_plane, n):
            hi = tuple(l + s for l, s in zip(lo, size))
            got = {c: fetch_box(c, lo, hi) for c in COMPONENTS}
            gd = None
            if lo[0] == mon_plane:
                gd = {c: fetch_dft(c, lo[1:], hi[1:]) for c in ("Ey", "Hz")}
            if got["Ex"] is None:
                continue                                   # not rank 0
            want, wd = oracle_crop(planes, dims, dt, spacing, lo, size, amp, ph, src_plane, mon_plane, n, coef_fn, src_profile)
            for c in COMPONENTS:
                worst = max(worst, rel_l2(got[c], want[c]))
                exact = exact and bool(np.array_equal(got[c], want[c]))
            if wd is not None and gd is not None and gd["Ey"] is not None:
                worst_dft = max(rel_l2(gd[c], wd[c]) for c in wd)
            out["crops"].append({"lo": list(lo), "size": list(size), "has_source_plane": bool(lo[0] <= src_plane < hi[0]),
                                 "at_monitor_plane": bool(lo[0] == mon_plane)})
        out["crop_rel_l2"], out["dft_rel_l2"], out["crop_bit_exact"] = worst, worst_dft, exact
    cs = fetch_checksums()
    if cs is not None:
        out["fields_sha"] = sha_of_checksums(cs)
    tol = 1e-4 if dtype == "float32" else 1e-10
    out["tolerance"] = tol
    if do_oracle and out["crops"]:
        out["ok"] = bool(worst <= tol and (worst_dft is None or worst_dft <= tol))
    return out
