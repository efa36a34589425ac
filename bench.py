#!/usr/bin/env python
"""bench.py — FDTD cell-updates/s of the B200 engine (and of the reference CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c2|c3|NxMxL]

One "step" = one full FDTD time step (H pass, E pass, sources, monitors) of the named workload.
Default workload "c4" = BASELINE.json configs[3]: 3-D 1024^3 vacuum, TFSF plane-wave source, DFT field-monitor
plane (5 frequencies), fp32 — the configuration the metric "cell-updates/s at 1/2/4/8 B200; HBM GB/s vs peak"
is quoted on (it fits one B200: 25.8 GB x 2 for the ping-pong sets).  Strong scaling: the same 1024^3 grid is
slab-decomposed along x over N ranks (one process per GPU, torchrun).  configs[1] (121^3) is L2-resident, so
its HBM fraction is meaningless; run it with --workload c2.

Prints ONE JSON line (rank 0).  Keys: see the driver contract; additionally
  roofline      fused-step (or H/E two-pass) kernel time from CUDA events on the engine stream, 48 B/cell fp32
  cpu_baseline  the oracle port of the reference NumPy path timed on this box's host cores (bounded sample)
  e2e           the same metric through the host-buffer API: upload of all six fields from pinned host
                memory + K steps + download of fields and monitor results, all inside the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C0 = 299792458.0
F0 = 193.4e12
COMPONENTS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
BYTES_PER_CELL = {"float32": 48, "float64": 96}          # 6 fields read once + written once


def parse_workload(name):
    name = name.lower()
    presets = {"c2": (121, 121, 121), "c3": (512, 512, 256), "c4": (1024, 1024, 1024), "c5": (2048, 1024, 512)}
    if name in presets:
        return name, presets[name]
    dims = tuple(int(v) for v in name.split("x"))
    assert len(dims) == 3
    return name, dims


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 5 ms from a thread
    (nvidia-smi -lms takes longer to start than an 8-GPU timed region lasts); nvidia-smi as a fallback."""

    def __init__(self, device=0):
        self.device, self.rows, self.stop = device, [], threading.Event()
        self.t = None

    def _nvml(self):
        import pynvml as N

        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.device)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self.stop.is_set():
            try:
                r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append((N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), mx, [k for k, b in bits.items() if r & b]))
            time.sleep(0.005)

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((float(out[0]), float(out[1]), [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")]))
            except Exception:
                return

    def _run(self):
        try:
            self._nvml()
        except Exception:
            self._smi()

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        time.sleep(0.02)
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({r for row in self.rows for r in row[2]})
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(max(r[1] for r in self.rows)),
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------
def workload_ops(dims, dt, spacing, x0=0, nxl=None):
    """TFSF +x plane source (whole plane, like the reference) at x = nx/4 and a DFT monitor plane at
    x = 3nx/4 (Ey, Hz, 5 frequencies), expressed as device ops clipped to the local slab [x0, x0+nxl)."""
    import prismo_b200 as pb

    nx, ny, nz = dims
    nxl = nx if nxl is None else nxl
    src_plane, mon_plane = nx // 4, (3 * nx) // 4
    src, mon = [], []
    last = x0 + nxl == nx

    def shape(c):
        n = [nxl, ny, nz]
        for ax in pb.grid.SHORT_AXES[c]:
            if ax == 0 and not last:
                continue
            n[ax] -= 1
        return n

    if x0 <= src_plane < x0 + nxl:
        i = src_plane - x0
        for c, tab in (("Ey", 0), ("Hz", 1)):
            s = shape(c)
            src.append(pb.SourceOp(c, (i, 0, 0), (i + 1, s[1], s[2]), tab))
    elif x0 + nxl <= src_plane < x0 + nxl + 3:
        # the two-step sweep recomputes the intermediate step on our ghost planes: it needs the right
        # neighbour's injection there too (ghost op, applied inside the sweep only)
        i = src_plane - x0
        for c, tab in (("Ey", 0), ("Hz", 1)):
            s = shape(c)
            src.append(pb.SourceOp(c, (i, 0, 0), (i + 1, s[1], s[2]), tab, ghost=True))
    if x0 <= mon_plane < x0 + nxl:
        i = mon_plane - x0
        for c in ("Ey", "Hz"):
            s = shape(c)
            mon.append(pb.MonitorOp(c, (i, 0, 0), (i + 1, s[1], s[2]), False, 5, 0))
    return src, mon


def tables(n_steps, dt, t0=0.0):
    """Host-evaluated per-step tables, the way the lowering does it (accumulated time)."""
    eta0 = np.sqrt((4 * np.pi * 1e-7) / 8.854187817e-12)
    w = 2 * np.pi * F0
    amp = np.zeros((n_steps, 2))
    ph = np.zeros((n_steps, 5), dtype=np.complex128)
    freqs = F0 * np.linspace(0.9, 1.1, 5)
    t = t0
    for s in range(n_steps):
        t += dt
        amp[s, 0] = -np.sin(w * t)                              # E[i_min,:] -= w(t)        (tfsf.py:333-338)
        amp[s, 1] = -(np.sin(w * (t - 0.5 * dt)) / eta0)        # H[i_min,:] -= w(t-dt/2)/eta0
        ph[s] = np.exp(-1j * (2 * np.pi * freqs) * t)
    return amp, ph, t


def workload_timestep():
    spacing = (2e-8, 2e-8, 2e-8)
    return 0.9 / (C0 * np.sqrt(3.0 / spacing[0] ** 2)), spacing


def make_engine(dims, dtype, rank=0, world=1, device=0, flags=0, span=None):
    """span = (x0, n) overrides the uniform x-slab of this rank (load-balanced slabs, bench_multi.py)."""
    import prismo_b200 as pb

    nx, ny, nz = dims
    dt, spacing = workload_timestep()
    base, rem = divmod(nx, world)
    nxl = base + (1 if rank < rem else 0)
    x0 = rank * base + min(rank, rem)
    if span is not None:
        x0, nxl = span
    eng = pb.Engine(3, (nxl, ny, nz), spacing, dt, dtype=dtype, device=device, nx_global=nx, x_offset=x0, flags=flags)
    return eng, dt, spacing, x0, nxl


class Medium:
    """Heterogeneous media of the named configs, x-invariant cross-sections on the (ny, nz) plane.
      c3  one Si ridge (eps 12.11) on an SiO2 (2.07) half space, air above; one Lorentz pole on the core.
      c5  directional coupler: two such ridges a gap apart; a gold pad (Drude pole, eps_inf = 1) beside them over the
          middle quarter of the length.
      aniso: the SiO2 half space becomes a uniaxial cladding, eps = diag(2.07, 2.07, 2.16) (optic axis z) — the named
          config 5's "anisotropic cladding" — applied as per-component Cb inside the E stage (opt-in extension)."""

    CLAD = (2.07, 2.07, 2.16)

    def __init__(self, name, dims, aniso=False):
        self.name, self.dims, self.aniso = name, dims, bool(aniso)
        nx, ny, nz = dims
        w, h = ny // 16, max(nz // 12, 2)
        if name == "c5":
            gap = max(w // 2, 2)
            self.cores = [(ny // 2 - gap // 2 - w, ny // 2 - gap // 2, nz // 2, nz // 2 + h),
                          (ny // 2 + gap - gap // 2, ny // 2 + gap - gap // 2 + w, nz // 2, nz // 2 + h)]
            self.pad = ((3 * nx) // 8, (5 * nx) // 8, ny // 2 + ny // 8, ny // 2 + ny // 8 + w, nz // 2, nz // 2 + max(h // 2, 1))
        else:
            self.cores = [(ny // 2 - ny // 16, ny // 2 + ny // 16, nz // 2, nz // 2 + h)]
            self.pad = None

    def eps(self, comp=0):
        nx, ny, nz = self.dims
        eps = np.ones((ny, nz), dtype=np.float64)
        eps[:, : nz // 2] = self.CLAD[comp] if self.aniso else 2.07
        for j0, j1, k0, k1 in self.cores:
            eps[j0:j1, k0:k1] = 12.11
        return eps

    def coefficients(self, dt, x_planes, jk=None):
        """Ca, Cb, Da, Db (lossless: Ca = Da = 1) of x_planes planes, optionally restricted to a (j, k) window.  With
        aniso, Cb is the tuple (Cb_x, Cb_y, Cb_z).  Same operations as core/solver.py:119-130 with sigma = 0."""
        eps0, mu0 = 8.854187817e-12, 4 * np.pi * 1e-7

        def cb(comp):
            eps = self.eps(comp)
            if jk is not None:
                eps = eps[jk[0]:jk[1], jk[2]:jk[3]]
            shp = (x_planes,) + eps.shape
            return np.ascontiguousarray(np.broadcast_to((dt / (eps0 * eps)) / (1 + 0.0 * eps), shp)), shp

        cbx, shp = cb(0)
        one = np.ones(shp)
        Cb = (cbx, cb(1)[0], cb(2)[0]) if self.aniso else cbx
        return one, Cb, one, one * ((dt / (mu0 * 1.0)) / 1.0)

    def shapes(self, spacing):
        """The same medium as a SHAPE LIST for the device rasteriser (geometry/shapes.py Box objects on cell coordinates
        i * d): an index range [a0, a1) becomes a box centred on it with faces a quarter cell beyond the end cells."""
        from prismo_b200 import geometry as G

        nx, ny, nz = self.dims

        def box(mat, r):
            c = [0.5 * (a0 + a1 - 1) * d for (a0, a1), d in zip(r, spacing)]
            size = [(a1 - a0 - 0.5) * d for (a0, a1), d in zip(r, spacing)]
            return G.Box(mat, c, size)

        clad = G.Material("cladding", self.CLAD if self.aniso else 2.07)
        out = [box(clad, ((-1, nx + 2), (0, ny), (0, nz // 2)))]
        for j0, j1, k0, k1 in self.cores:
            out.append(box(G.Material("Si", 12.11), ((-1, nx + 2), (j0, j1), (k0, k1))))
        return out

    def mode_profiles(self):
        """Transverse profile of the mode source: a Gaussian centred on the (first) Si core — it stands in for the solved
        mode profile a ModeSource injects (sources/mode.py:255-361) — per injected component on its own (ny', nz') grid."""
        nx, ny, nz = self.dims
        j0, j1, k0, k1 = self.cores[0]
        y = (np.arange(ny) - 0.5 * (j0 + j1)) / max(0.5 * (j1 - j0), 1.0)
        z = (np.arange(nz - 1) - 0.5 * (k0 + k1)) / max(0.5 * (k1 - k0), 1.0)
        prof = np.exp(-(y[:, None] ** 2) - (z[None, :] ** 2))
        return {"Ey": prof, "Hz": prof}

    def ade_ops(self, dt, x0=0, nxl=None):
        """Recursions clipped to planes [x0, x0 + nxl).  c3: one Lorentz pole (resonance at 1.2 um) on the core; c5: the same
        on both cores plus a Drude pole (gold: omega_p = 1.37e16 rad/s, gamma = 4.05e13 1/s) on the pad — coefficients of
        materials/dispersion.py:189-231 and :267-286, recursions of materials/ade.py:116-149."""
        from prismo_b200.engine import AdeOp

        nx, ny, nz = self.dims
        nxl = nx if nxl is None else nxl
        w0, de, gam = 2 * np.pi * C0 / 1.2e-6, 1.0, 1e13
        den = 4.0 + 2 * gam * dt + w0 ** 2 * dt ** 2
        c0 = 2 * de * w0 ** 2 * dt ** 2 / den
        c2, c3 = (8.0 - 2 * w0 ** 2 * dt ** 2) / den, -(4.0 - 2 * gam * dt + w0 ** 2 * dt ** 2) / den
        ops = []
        shorts = (("Ex", (0, 1, 1)), ("Ey", (1, 0, 1)), ("Ez", (1, 1, 0)))

        def clip(c, short, a, b):
            planes = nxl - (1 if (short[0] and x0 + nxl == nx) else 0)        # this component's local x extent
            return max(a - x0, 0), min(b - x0, planes)

        for j0, j1, k0, k1 in self.cores:
            for c, short in shorts:
                a, b = clip(c, short, 0, nx)
                if b > a:
                    ops.append(AdeOp(c, 0, (a, j0, k0), (b, j1 - short[1], k1 - short[2]), c0, c0, c2, c3))
        if self.pad is not None:
            i0, i1, j0, j1, k0, k1 = self.pad
            e = np.exp(-4.05e13 * dt)
            d0 = 1.37e16 ** 2 / 4.05e13 * (1.0 - e)
            for c, short in shorts:
                a, b = clip(c, short, i0, i1)
                if b > a:
                    ops.append(AdeOp(c, 1, (a, j0, k0), (b, j1 - short[1], k1 - short[2]), d0, e))
        return ops


def workload_text(name, dims):
    if name == "c5":
        return (f"c5: 3-D {dims[0]}x{dims[1]}x{dims[2]} directional coupler (two Si ridges on SiO2, cell-centred Ca,Cb,Da,Db "
                "arrays), Lorentz pole on both cores + Au Drude pad (9 recursions), profiled mode-source plane, 2 ports x 2 "
                "planes x 6 components x 3-frequency DFT + FieldMonitor DFT plane; mode-overlap S-parameters reduced on the "
                "device; uniaxial cladding eps = diag(2.07, 2.07, 2.16) as per-component Cb in the E stage unless --no-aniso")
    return (f"{name}: 3-D {dims[0]}x{dims[1]}x{dims[2]} Si ridge on SiO2 (cell-centred Ca,Cb,Da,Db arrays), Lorentz pole on "
            "the core (3 recursions, in-sweep), profiled mode-source plane, FieldMonitor DFT plane (Ey,Hz x 5 freq)")


def port_monitor_ops(dims, x0=0, nxl=None, n_freq=3):
    """config 5: two ports, two planes each (8 planes apart), all six components on the box common to them, n_freq running
    DFTs each — what a pair of region-correct FluxMonitors per port lowers to (prismo_b200/lowering.py:_FluxRegionBinder).
    Returns [(port, which, global plane, [MonitorOp x 6 clipped to the slab or None])]."""
    import prismo_b200 as pb

    nx, ny, nz = dims
    nxl = nx if nxl is None else nxl
    planes = [(0, 0, nx // 8), (0, 1, nx // 8 + 8), (1, 0, (7 * nx) // 8 - 8), (1, 1, (7 * nx) // 8)]
    out = []
    for port, which, p in planes:
        ops = None
        if x0 <= p < x0 + nxl:
            i = p - x0
            ops = [pb.MonitorOp(c, (i, 0, 0), (i + 1, ny - 1, nz - 1), False, n_freq, 0) for c in COMPONENTS]
        out.append((port, which, p, ops))
    return out


def install_workload(eng, name, dims, dt, spacing, args, x0=0, nxl=None):
    """Coefficients, sources, monitors and recursions of the named workload on this engine (one GPU or one x-slab).
    Returns a dict: mon_ids (the Ey / Hz DFT plane of the self-check), ports, src_profile, coef_fn, medium, real."""
    import prismo_b200 as pb

    nx = dims[0]
    nxl = nx if nxl is None else nxl
    real = name in ("c3", "c5") and not args.no_ops and not args.physics     # the BASELINE configs as named
    if real:
        args.het = True
    aniso = bool(args.het and getattr(args, "aniso", False))
    med = Medium(name, dims, aniso) if args.het else None
    setup = None
    if med is not None:
        planes = min(nxl + 1, nx - x0)                                      # + the right neighbour's first plane on a slab
        t0 = time.perf_counter()
        if getattr(args, "host_coeffs", False):
            Ca, Cb, Da, Db = med.coefficients(dt, planes)
            if aniso:
                eng.set_coeffs_aniso(Ca, *Cb, Da, Db)
            else:
                eng.set_coeffs(Ca, Cb, Da, Db)
            how = "host-painted fp64 arrays uploaded (fdtd_set_coeffs%s)" % ("_aniso" if aniso else "")
        else:
            ax = [np.arange(n) * d for n, d in zip(dims, spacing)]
            eng.set_option("het_indexed", int(not getattr(args, "no_indexed", False)))
            eng.rasterize(med.shapes(spacing), ax[0][x0:x0 + planes], ax[1], ax[2])
            how = "shape list rasterised on the device (fdtd_rasterize: %d boxes -> %d coefficient arrays)" % (
                1 + len(med.cores), 6 if aniso else 4)
        eng.sync()
        setup = {"coefficients": how, "seconds": time.perf_counter() - t0, "aniso": aniso,
                 "index_coded": not (getattr(args, "no_indexed", False) or getattr(args, "host_coeffs", False))}
    src, mon = workload_ops(dims, dt, spacing, x0, nxl)
    src_profile, ports = None, []
    if real:
        src_profile = med.mode_profiles()
        # (no ghost copies: heterogeneous media step one sweep at a time, which never recomputes a neighbour's planes)
        src = [pb.SourceOp(o.component, o.lo, o.hi, o.table, src_profile[o.component][None, :, :]) for o in src if not o.ghost]
        for op in med.ade_ops(dt, x0, nxl):
            eng.add_ade_op(op)
    if args.no_ops:
        src, mon = [], []
    for op in src:
        eng.add_source_op(op)
    mon_ids = [eng.add_monitor_op(op) for op in mon]
    if real and name == "c5":
        for port, which, p, ops in port_monitor_ops(dims, x0, nxl):
            ports.append((port, which, p, None if ops is None else [eng.add_monitor_op(o) for o in ops]))
    coef_fn = None
    if med is not None:
        coef_fn = lambda lo, sd: med.coefficients(dt, sd[0], (lo[1], lo[1] + sd[1], lo[2], lo[2] + sd[2]))  # noqa: E731
    return {"mon_ids": mon_ids, "mon": mon, "ports": ports, "src_profile": src_profile, "coef_fn": coef_fn, "medium": med,
            "real": real, "setup": setup, "aniso": aniso}


def port_s_parameters(eng, ports, med, dims, spacing, gather=None):
    """2-port S-parameters of config 5 from the four resident DFT planes: mode-overlap sums on the device
    (fdtd_mode_overlap), forward / backward split between the two planes of a port, S11 = b1/a1, S21 = a2/a1
    (prismo_b200/postprocess.py; utils/mode_matching.py:41-131, :228-293; analysis/sparameters.py:88-122).
    gather: callable merging {(port, which): coefficients} over ranks (x-slabs), None on one GPU."""
    from types import SimpleNamespace

    from prismo_b200 import postprocess as PP

    nx, ny, nz = dims
    prof = med.mode_profiles()["Ey"][: ny - 1, : nz - 1]
    zero = np.zeros_like(prof)
    eta = 377.0 / np.sqrt(12.11)
    mode = SimpleNamespace(Ex=zero, Ey=prof, Ez=zero, Hx=zero, Hy=zero, Hz=prof / eta)      # +x-going quasi-TE profile
    power = PP.mode_power(mode, "x", spacing[1], spacing[2])
    mine = {}
    for port, which, p, ids in ports:
        if ids is not None:
            num = eng.mode_overlap(ids, "x", [getattr(mode, c)[None, :, :] for c in COMPONENTS])
            mine[(port, which)] = num * spacing[1] * spacing[2] / power
    coeffs = gather(mine) if gather else mine
    if coeffs is None or len(coeffs) < 4:
        return None
    lams = C0 / (F0 * np.linspace(0.9, 1.1, 5)[:3])
    s11, s21 = [], []
    for i, lam in enumerate(lams):
        a1, b1 = PP.separate_forward_backward(coeffs[(0, 0)][i], coeffs[(0, 1)][i], 2.4, 8 * spacing[0], lam)
        a2, _ = PP.separate_forward_backward(coeffs[(1, 0)][i], coeffs[(1, 1)][i], 2.4, 8 * spacing[0], lam)
        s11.append(b1 / a1)
        s21.append(a2 / a1)
    fmt = lambda v: [[float(np.real(x)), float(np.imag(x))] for x in v]  # noqa: E731
    return {"wavelengths_um": [float(x * 1e6) for x in lams], "S11": fmt(s11), "S21": fmt(s21),
            "note": "mode-overlap sums reduced on the device over resident DFT planes; the reference's update scheme is "
                    "unstable (SURVEY F4), so the values exercise the extraction path, not a physical device"}


def seed_fields(eng, dims, x0=0):
    """Small white-noise fields as a pure function of the GLOBAL cell index (bench_check.py), so every decomposition
    of the grid starts from the same global state and the end-of-run checksums are comparable across N."""
    import bench_check as BC

    return BC.seed_fields(eng, dims, x0)


def fields_sha(eng):
    import bench_check as BC

    return BC.sha_of_checksums({c: eng.plane_checksums(c) for c in COMPONENTS})


def self_check(eng, dims, dt, spacing, args, mon_ids, do_oracle, coef_fn=None, src_profile=None):
    """Re-seed, run a few steps, compare two crops and the DFT plane with the oracle, checksum the fields (bench_check.py)."""
    import bench_check as BC

    def reseed(n, amp, ph):
        eng.zero_fields()
        planes = BC.seed_fields(eng, dims)
        for i in mon_ids:
            op = eng._mon_ops[i]
            eng.set_dft(i, np.zeros((op.n_freq,) + op.shape, dtype=np.complex128))
        eng.set_tables(n, amp, ph)
        return planes

    def fetch_dft(c, lo2, hi2):
        i = mon_ids[("Ey", "Hz").index(c)]
        return eng.dft(i)[:, 0, lo2[0]:hi2[0], lo2[1]:hi2[1]]

    return BC.run_check(dims, dt, spacing, args.dtype, tables, dims[0] // 4, (3 * dims[0]) // 4, reseed, eng.run,
                        lambda c, lo, hi: eng.download_box(c, lo, hi), fetch_dft,
                        lambda: {c: eng.plane_checksums(c) for c in COMPONENTS}, do_oracle=do_oracle and bool(mon_ids),
                        coef_fn=coef_fn, src_profile=src_profile)


# ---------------------------------------------------------------------------------------------------------
def cpu_arm(side):
    """The CPU implementation of the path on a side^3 sample of the c4 workload (vacuum, TFSF plane, 5-frequency DFT
    plane): -> (step callable, kind, description).  kind "reference" = the UNMODIFIED rithulkamesh/prismo package on its
    stock NumPy backend (Simulation.step, core/simulation.py:147-164), imported from oracle/_ref — the staged copy that
    travels to the GPU box (oracle/stage_reference.py) — or from /root/reference/src in the build container; kind "port" =
    the oracle restatement (oracle/sim.py: the same NumPy expressions one for one) where no copy of the reference exists.
    FDTD_B200_CPU_ARM=port forces the port."""
    res, pml = 50e6, 10
    size = ((side - 2 * pml - 0.5) / res,) * 3
    L = size[0]
    src_kw = dict(center=(L / 2, L / 2, L / 2), size=(L / 2, L / 2, L / 2), direction="+x", polarization="y", frequency=F0,
                  pulse=False)
    mon_kw = dict(center=(0.75 * L, L / 2, L / 2), size=(0.0, L, L), components=["Ey", "Hz"], time_domain=False,
                  frequencies=list(F0 * np.linspace(0.9, 1.1, 5)))
    why = "forced by FDTD_B200_CPU_ARM=port"
    if os.environ.get("FDTD_B200_CPU_ARM", "reference") != "port":
        try:
            from oracle import ref_loader

            if not ref_loader.available():
                raise ImportError(f"no reference package at {ref_loader.REF_SRC}")
            prismo = ref_loader.load()
            import prismo.monitors.field
            import prismo.sources.tfsf

            prismo.set_backend("numpy")
            sim = prismo.Simulation(size=size, resolution=res, pml_layers=pml, courant_factor=0.9)
            if tuple(sim.grid.dimensions) != (side, side, side):
                raise ValueError(f"reference grid {sim.grid.dimensions} != {side}^3")
            sim.add_source(prismo.sources.tfsf.TFSFSource(**src_kw))
            sim.add_monitor(prismo.monitors.field.FieldMonitor(**mon_kw))
            return sim.step, "reference", (f"unmodified prismo {getattr(prismo, '__version__', '?')} (NumPy backend, "
                                           f"Simulation.step) imported from {os.path.relpath(ref_loader.REF_SRC, ROOT)}")
        except Exception as ex:                 # no staged copy / import problem: the port below
            why = f"reference not importable ({type(ex).__name__}: {ex})"
    from oracle import sim as O

    s = O.OSimulation(size, res, pml, 0.9)
    assert s.grid.dims == (side, side, side), s.grid.dims
    s.add_source(O.TFSFSource(src_kw["center"], src_kw["size"], "+x", "y", F0, pulse=False))
    s.add_monitor(O.FieldMonitor(mon_kw["center"], mon_kw["size"], components=["Ey", "Hz"], time_domain=False,
                                 frequencies=mon_kw["frequencies"]))
    return s.step, "port", f"oracle port of the reference NumPy path (oracle/sim.py); {why}"


def cpu_baseline_sample(dims, seconds_budget=20.0, steps=3):
    """The reference's CPU path (cpu_arm) on a bounded cubic sample of the workload."""
    cells_per_s = 6.0e6                       # survey estimate, only used to size the sample
    side = int(min(min(dims), max(48, round((cells_per_s * seconds_budget / (steps + 1)) ** (1 / 3)))))
    side = min(side, 224)
    step, kind, what = cpu_arm(side)
    step()                                    # warm-up (page faults, first-touch)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return side ** 3 * steps / dt, side, steps, dt, kind, what


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (cpu_arm): the
    unmodified package from the staged copy where it exists, the oracle port otherwise.  NumPy's elementwise kernels are
    single-threaded, so one core is all it can use."""
    if rank != 0:
        return
    name, dims = parse_workload(args.workload)
    total = args.steps + args.warmup
    side = int(min(min(dims), max(48, round((6.0e6 * 120.0 / max(total, 1)) ** (1 / 3)))))
    side = min(side, 224)
    step, kind, what = cpu_arm(side)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = side ** 3 * args.steps / el
    sample = (f"{side}^3 sub-grid of the {name} workload (vacuum, TFSF plane, 5-frequency DFT plane), "
              f"{args.steps} steps, NumPy {np.__version__}, fp64; {what}")
    line = {"impl": "reference", "metric": "fdtd_cell_updates_per_s", "value": value, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: 3-D {dims[0]}x{dims[1]}x{dims[2]} vacuum + TFSF + DFT plane "
                                   f"(timed on a bounded {side}^3 sample)", "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": kind, "sample": sample,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def run_c1(args, local):
    """configs[0]: 2-D 1000x1000 Si strip waveguide (eps 12.11 in 2.07), Gaussian-beam line source, DFTMonitor corner
    patch (reference placeholder semantics) + FieldMonitor DFT line, 11 frequencies.  Launch-bound (1e6 cells):
    timed through the CUDA-graph step loop.  Parity for this shape: tests (src2d_gauss, mon2d_all, upd2d_het)."""
    import prismo_b200 as pb

    nx = ny = 1000
    d = 2e-8
    dt = 0.9 / (C0 * np.sqrt(2.0 / d ** 2))
    eng = pb.Engine(2, (nx, ny, 1), (d, d, 0.0), dt, dtype=args.dtype, device=local)
    eps = np.full((nx, ny), 2.07)
    eps[:, ny // 2 - 11: ny // 2 + 11] = 12.11
    eps0, mu0 = 8.854187817e-12, 4 * np.pi * 1e-7
    one = np.ones((nx, ny))
    eng.set_coeffs(one, dt / (eps0 * eps), one, one * (dt / mu0))
    y = (np.arange(400, 600) - 500) * d
    prof = np.exp(-(y / 1e-6) ** 2)[None, :]
    eng.add_source_op(pb.SourceOp("Ey", (60, 400), (61, 600), 0, prof))
    eng.add_source_op(pb.SourceOp("Hz", (60, 400), (61, 600), 0, prof, 377.0))
    for c in ("Ex", "Ey", "Ez"):
        eng.add_monitor_op(pb.MonitorOp(c, (0, 0), (10, 10), False, 11, 0))
    eng.add_monitor_op(pb.MonitorOp("Ey", (800, 400), (801, 600), False, 11, 0))
    total = args.warmup + args.steps
    t = (np.arange(total) + 1) * dt
    amp = (np.exp(-0.5 * ((t - 3e-14) / 1e-14) ** 2) * np.sin(2 * np.pi * F0 * t))[:, None]
    ph = np.exp(-1j * 2 * np.pi * (C0 / np.linspace(1.5e-6, 1.6e-6, 11))[None, :] * t[:, None])
    eng.set_tables(total, amp, ph)
    eng.run(args.warmup)
    eng.sync()
    l0 = eng.kernel_launches
    with ClockSampler(local) as clk:
        eng.timer_start()
        eng.run(args.steps)
        ms = eng.timer_stop()
    cells = nx * ny
    value = cells * args.steps / (ms * 1e-3)
    bpc = (48 if args.dtype == "float32" else 96) // 1
    line = {"metric": "fdtd_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
            "config": {"workload": "c1: 2-D 1000x1000 Si strip waveguide, Gaussian-beam line source, DFT monitors (11 freq)",
                       "l2": "working set 24-48 MB: L2-resident, launch-bound (HBM fraction not meaningful)",
                       "kernel_path": "k_h2d + k_e2d + k_sources + k_monitors, 16 steps per CUDA graph replay"},
            "roofline": {"bound": "hbm", "achieved": bpc * value / 1e9, "peak": peaks()[0], "unit": "GB/s",
                         "frac": bpc * value / 1e9 / peaks()[0], "traffic": None,
                         "note": "algorithmic GB/s of an L2-resident grid"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": eng.kernel_launches - l0, "clocks": clk.summary()}
    print(json.dumps(line), flush=True)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--two-pass", action="store_true", help="force the un-fused H/E kernels")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fast-f64", action="store_true", help="fp64 with folded FMA arithmetic (within 1e-10, not bit-exact)")
    ap.add_argument("--het", action="store_true", help="heterogeneous medium: Si ridge (eps 12.11) on SiO2 (2.07), "
                                                       "cell-centred coefficient arrays (64 B/cell)")
    ap.add_argument("--physics", action="store_true", help="opt-in physics mode: stable Yee leap-frog + 10-cell CPML "
                    "(two-pass kernels with slab psi updates; no reference numbers exist for it)")
    ap.add_argument("--no-ops", action="store_true", help="bare field update: no source, no monitor (tuning only)")
    ap.add_argument("--no-check", action="store_true", help="skip the post-run self-check against the oracle")
    ap.add_argument("--aniso", action="store_true", help="heterogeneous workloads: uniaxial cladding eps = diag(2.07, 2.07, "
                    "2.16) as per-component Cb inside the E stage (default for c5, whose named config has it)")
    ap.add_argument("--no-aniso", action="store_true", help="c5 without its anisotropic cladding (one Cb per cell, 64 B/cell)")
    ap.add_argument("--no-indexed", action="store_true", help="heterogeneous sweep streams the 4 (6) coefficient arrays instead "
                    "of one material-index byte per cell + a shared-memory material table (the default with device-painted "
                    "coefficients)")
    ap.add_argument("--host-coeffs", action="store_true", help="paint Ca..Db on the host and upload them instead of "
                    "rasterising the shape list on the device")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload.lower() == "c5" and not args.no_aniso and not args.physics and not args.no_ops:
        args.aniso = True

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if world > 1:
        from bench_multi import run_multi            # slab-decomposed path (torch.distributed / NCCL)

        run_multi(args, rank, world, local)
        return

    import prismo_b200 as pb
    from prismo_b200 import _lib

    if args.workload.lower() == "c1":
        return run_c1(args, local)
    name, dims = parse_workload(args.workload)
    cells = dims[0] * dims[1] * dims[2]
    flags = (_lib.FLAG_TWO_PASS if args.two_pass else 0) | (_lib.FLAG_FAST_F64 if args.fast_f64 else 0)
    if args.physics:
        flags |= _lib.FLAG_YEE
    eng, dt, spacing, x0, nxl = make_engine(dims, args.dtype, device=local, flags=flags)
    if args.physics:
        from prismo_b200 import cpml

        thick = int(os.environ.get("FDTD_B200_BENCH_CPML", "10"))       # 0: tuning experiment (no absorbing layer)
        eng.set_cpml(thick, cpml.coefficient_table(dims, spacing, dt, cpml.PMLParams(thickness=max(thick, 1))))
    wl = install_workload(eng, name, dims, dt, spacing, args)
    mon_ids, src_profile, real_c3 = wl["mon_ids"], wl["src_profile"], wl["real"]
    total_steps = args.warmup + args.steps
    amp, ph, _ = tables(total_steps, dt)
    eng.set_tables(total_steps, amp, ph)
    seed_fields(eng, dims)

    # ---- timed region: W warm-up steps, then exactly K steps between events on the engine stream -----------
    eng.run(args.warmup)
    eng.sync()
    l0 = eng.kernel_launches
    small = cells * BYTES_PER_CELL[args.dtype] < 400e6          # L2-resident grids are launch-bound: time the graph loop
    with ClockSampler(local) as clk:
        if small:
            eng.timer_start()
            eng.run(args.steps)
            t_ms = eng.timer_stop()
            prof = {"total_ms": t_ms, "h_or_fused_ms": t_ms, "e_ms": 0.0, "post_ms": 0.0}
        else:
            prof = eng.run_profiled(args.steps)
    launches = eng.kernel_launches - l0
    timed_sha = fields_sha(eng)                  # state after W + K steps: comparable across N for equal W, K
    s_params = port_s_parameters(eng, wl["ports"], wl["medium"], dims, spacing) if wl["ports"] else None
    ms = prof["total_ms"]
    value = cells * args.steps / (ms * 1e-3)

    # dominant kernel(s): fused sweep (one launch per step) or H + E pass
    kern_ms = prof["h_or_fused_ms"] + prof["e_ms"]
    peak, peak_src = peaks()
    indexed = bool(wl["setup"] and wl["setup"].get("index_coded"))
    n_coef = (6 if wl["aniso"] else 4) * int(args.het)                     # + Ca,Cb,Da,Db (+ Cb_y, Cb_z) reads
    bpc = BYTES_PER_CELL[args.dtype] + (4 if args.dtype == "float32" else 8) * n_coef
    achieved = bpc * cells * args.steps / (kern_ms * 1e-3) / 1e9
    fused = not (args.two_pass or args.physics)
    tb2 = fused and os.environ.get("FDTD_B200_TB2", "1") != "0" and args.steps >= 2 and not args.het
    kname = ("k_fused3d_tb2x (TMA-fed, 1 launch per TWO steps)" if tb2 else
             ("k_fused3d_het" + ("<ADE> (dispersive recursions applied in-sweep)" if real_c3 else "") + " (1 launch/step, "
              + (f"6 field arrays + 1 material-index byte per cell read (material table of {n_coef} values per entry in shared "
                 "memory), 6 written)" if indexed else f"6 field + {n_coef} coefficient arrays read, 6 written)")) if args.het else "k_fused3d (1 launch/step)") if fused \
        else ({"2": "k_fused3d_yeex (physics mode: Yee leap-frog + CPML slabs in ONE TMA-fed sweep per step, psi ping-pong)",
               "1": "k_fused3d_yee (physics mode: Yee leap-frog + CPML slabs fused into ONE sweep per step, psi ping-pong)",
               "0": "k_h3d_yee + k_e3d_yee (physics mode: Yee leap-frog + CPML slabs, 2 launches/step)"}[
                   os.environ.get("FDTD_B200_YEE_FUSED", "2")] if args.physics
              else "k_h3d + k_e3d (2 launches/step)")
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        key = f"{name}:{args.dtype}:" + ("physics" if args.physics else ("two-step" if tb2 else "one-step"))
        if args.het:
            key = f"{name}:{args.dtype}:het" + ("-aniso" if wl["aniso"] else "") + ("-indexed" if indexed else "")
        if (fused or args.physics) and key in tr:
            traffic = tr[key]["dram_bytes"]          # bytes per launch, from the committed ncu capture of this kernel
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": kname,
                "note": ("achieved = ALGORITHMIC bytes (48 B per cell-update, SURVEY 8d) / kernel time; the two-step "
                         "sweep keeps the intermediate step on chip, so its real DRAM traffic is 28.2 B per cell-update "
                         "(ncu: profiles/r02_ncu_summary.md, 84 % of the copy peak) and frac can exceed 1") if tb2 else
                        (f"achieved = ALGORITHMIC bytes ({bpc} B per cell-update: fields + the coefficient values the update "
                         "consumes, SURVEY 8d) / kernel time; with material-index coding the sweep reads one index byte per "
                         f"cell instead of {n_coef} coefficient values, so its real DRAM traffic is about "
                         f"{BYTES_PER_CELL[args.dtype] + 1} B per cell-update and frac can exceed the DRAM fraction") if indexed else None,
                "algorithmic_bytes_per_launch": (2 if tb2 else 1) * bpc * cells if fused else bpc * cells / 2,
                "odd_last_step": "one-step sweep" if (tb2 and args.steps % 2) else None,
                "kernel_ms_per_step": kern_ms / args.steps, "post_ms_per_step": prof["post_ms"] / args.steps}

    # ---- e2e: host buffers in, host buffers out ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(eng, dims, args, mon_ids, dt)

    check = None
    if not args.no_check:
        check = self_check(eng, dims, dt, spacing, args, mon_ids, do_oracle=not (args.physics or args.no_ops),
                           coef_fn=wl["coef_fn"], src_profile=src_profile)
        check["timed_fields_sha"] = timed_sha

    cpu = None
    if not args.no_cpu:
        v, side, st, el, kind, what = cpu_baseline_sample(dims)
        cpu = {"value": v, "unit": "cell-updates/s", "cores": 1, "kind": kind, "host_cores_available": os.cpu_count(),
               "sample": f"{side}^3 sub-grid of the workload, {st} steps in {el:.1f} s, {what} (single-threaded "
                         f"elementwise ufuncs), fp64"}

    line = {"metric": "fdtd_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64",
            "data": "synthetic",
            "config": {"workload": workload_text(name, dims) if real_c3 else
                                   (f"{name}: 3-D {dims[0]}x{dims[1]}x{dims[2]} " + ("Si ridge on SiO2 (cell-centred coefficient arrays)"
                                    if args.het else "vacuum (uniform coefficients)") + ", TFSF +x "
                                    f"plane source, FieldMonitor DFT plane (Ey,Hz x 5 freq)"),
                       "l2": f"working set {2 * bpc // 2 * cells / 1e9:.1f} GB >> 126 MB L2 (no flush needed)"
                             if cells * bpc / 2 > 1e9 else "working set fits L2: HBM fraction not meaningful",
                       "parallelism": "1 GPU", "kernel_path": ("temporally blocked fused sweep (2 steps per HBM pass), ping-pong" if tb2 else
                                       "fused single sweep, ping-pong") if fused else
                                       (("physics mode (opt-in, parity unpinned): Yee leap-frog + 10-cell CPML, "
                                         + ("fused one-sweep step" if os.environ.get("FDTD_B200_YEE_FUSED", "2") != "0" else "two-pass"))
                                        if args.physics else "two-pass")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(), "check": check}
    if s_params is not None:
        line["s_params"] = s_params
    if wl["setup"] is not None:
        line["setup"] = wl["setup"]
    print(json.dumps(line), flush=True)
    eng.close()


def run_e2e(eng, dims, args, mon_ids, dt):
    """Whole job through the host-buffer API: pinned host fields -> device, K steps, results -> host."""
    import torch

    esz = 4 if args.dtype == "float32" else 8
    tdt = torch.float32 if args.dtype == "float32" else torch.float64
    host = {}
    for c in COMPONENTS:
        t = torch.zeros(eng.field_shape(c), dtype=tdt, pin_memory=True)
        host[c] = t.numpy()
        eng.download(c, host[c])                     # start the job from the current state
    amp, ph, _ = tables(args.steps, dt)
    h2d = sum(a.nbytes for a in host.values()) + amp.nbytes + ph.nbytes
    t0 = time.perf_counter()
    for c in COMPONENTS:
        eng.upload(c, host[c])
    eng.set_tables(args.steps, amp, ph)
    eng.run(args.steps)
    d2h = 0
    for c in COMPONENTS:
        eng.download(c, host[c])
        d2h += host[c].nbytes
    for i in mon_ids:
        d2h += eng.dft(i).nbytes
    el = time.perf_counter() - t0
    cells = dims[0] * dims[1] * dims[2]
    return {"value": cells * args.steps / el, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / args.steps,
            "d2h_bytes_per_step": d2h / args.steps, "seconds": el,
            "what": "upload 6 fields from pinned host memory + tables, K steps, download 6 fields + DFT planes"}


if __name__ == "__main__":
    main()
