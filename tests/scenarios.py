"""Shared test scenarios: one spec -> the same simulation built three ways.

  build_reference(spec, prismo)   the REAL reference objects (build container only)
  build_oracle(spec)              oracle.sim.OSimulation (NumPy restatement, travels to the GPU box)
  build_mirror(spec, pb)          prismo_b200's interface-compatible classes (the product)

Specs are plain data so the golden generator (tests/golden/make_golden.py) and every parity test agree
on inputs.  Initial fields are seeded white noise, E ~ N(0,1), H ~ N(0,1)/377 (SURVEY §8c conditioning
rule: the reference scheme is unstable, broadband noise keeps fp32/fp64 comparisons meaningful).
"""
from __future__ import annotations

import types

import numpy as np

COMPONENTS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
F0 = 193.4e12


def _wf(kind, **kw):
    return ("waveform", kind, kw)


def chirp(t):
    """User waveform function of the CustomWaveform scenario: a linear chirp under a raised-cosine ramp."""
    import math

    return math.sin(2 * math.pi * F0 * t * (1.0 + 2.0e13 * t)) * (0.5 - 0.5 * math.cos(min(t * 4e15, math.pi)))


_WAVEFORM_FUNCS = {"chirp": chirp}


def _wf_kwargs(kw):
    """Scenario specs name waveform functions by string; resolve them to the callables."""
    return {k: (_WAVEFORM_FUNCS[v] if k == "waveform_func" else v) for k, v in kw.items()}


SCENARIOS = {
    # ---- bare updates ---------------------------------------------------------------------------
    "upd3d_vac": dict(size=(0.9e-6, 0.7e-6, 0.6e-6), resolution=20e6, pml=3, courant=0.9, steps=12),
    "upd3d_het": dict(size=(0.8e-6, 0.6e-6, 0.5e-6), resolution=(20e6, 25e6, 30e6), pml=2, courant=0.5,
                      materials="random", steps=10),
    "upd3d_odd": dict(size=(1.65e-6, 0.35e-6, 1.05e-6), resolution=20e6, pml=1, courant=0.1,
                      materials="random", steps=6),
    "upd2d_vac": dict(size=(1.5e-6, 1.1e-6, 0.0), resolution=20e6, pml=4, courant=0.9, steps=12),
    "upd2d_het": dict(size=(1.2e-6, 0.9e-6, 0.0), resolution=(30e6, 20e6, 20e6), pml=3, courant=0.5,
                      materials="random", steps=10),
    # 2-D gates (solver.py:321,367): only Ez excited / only Ex,Ey excited / lossy magnetic medium
    "upd2d_gate_tm": dict(size=(1.0e-6, 0.8e-6, 0.0), resolution=20e6, pml=2, courant=0.5, materials="random",
                          init_only=("Ez", "Hz"), steps=6),
    "upd2d_gate_te": dict(size=(1.0e-6, 0.8e-6, 0.0), resolution=20e6, pml=2, courant=0.5, materials="random",
                          init_only=("Ex", "Hx", "Hy"), steps=6),
    # ---- sources ------------------------------------------------------------------------------------
    "src3d_point": dict(size=(0.6e-6, 0.6e-6, 0.6e-6), resolution=20e6, pml=2, courant=0.5, init="zero", steps=8,
                        sources=[("PointSource", dict(position=(0.3e-6, 0.25e-6, 0.2e-6), component="Ez",
                                                      waveform=_wf("GaussianPulse", frequency=F0, pulse_width=2e-16))),
                                 ("PointSource", dict(position=(0.1e-6, 0.1e-6, 0.4e-6), component="Hy",
                                                      waveform=_wf("RickerWavelet", frequency=F0))),
                                 ("ElectricDipole", dict(position=(0.3e-6, 0.25e-6, 0.2e-6), polarization="z",
                                                         frequency=F0, pulse=False, amplitude=0.5))]),
    # row a11: a user-supplied waveform function and the magnetic dipole constructors (sources/waveform.py:196-229,
    # sources/point.py:142-206), CW and pulsed
    "src3d_custom": dict(size=(0.6e-6, 0.5e-6, 0.4e-6), resolution=20e6, pml=2, courant=0.5, init="zero", steps=9,
                         sources=[("PointSource", dict(position=(0.25e-6, 0.2e-6, 0.2e-6), component="Ex",
                                                       waveform=_wf("CustomWaveform", waveform_func="chirp", amplitude=0.7))),
                                  ("MagneticDipole", dict(position=(0.3e-6, 0.25e-6, 0.15e-6), polarization="y",
                                                          frequency=F0, pulse=False, amplitude=2.0, phase=0.4)),
                                  ("MagneticDipole", dict(position=(0.1e-6, 0.3e-6, 0.25e-6), polarization="z",
                                                          frequency=F0, pulse=True, pulse_width=2.5e-16))]),
    "src3d_plane": dict(size=(0.6e-6, 0.5e-6, 0.4e-6), resolution=20e6, pml=2, courant=0.5, steps=8,
                        sources=[("PlaneWaveSource", dict(center=(0.2e-6, 0.25e-6, 0.2e-6), size=(0.0, 0.3e-6, 0.2e-6),
                                                          direction="-x", polarization="z", frequency=F0,
                                                          pulse=True, pulse_width=3e-16, amplitude=2.0, phase=0.3))]),
    "src3d_tfsf": dict(size=(0.6e-6, 0.5e-6, 0.4e-6), resolution=20e6, pml=3, courant=0.9, steps=8,
                       sources=[("TFSFSource", dict(center=(0.3e-6, 0.25e-6, 0.2e-6), size=(0.3e-6, 0.3e-6, 0.2e-6),
                                                    direction="+x", polarization="y", frequency=F0, pulse=False)),
                                ("TFSFSource", dict(center=(0.3e-6, 0.25e-6, 0.2e-6), size=(0.3e-6, 0.2e-6, 0.2e-6),
                                                    direction="-z", polarization="x", frequency=F0, pulse=True,
                                                    pulse_width=2e-16)),
                                ("TFSFSource", dict(center=(0.3e-6, 0.25e-6, 0.2e-6), size=(0.3e-6, 0.2e-6, 0.2e-6),
                                                    direction="y", polarization="z", frequency=F0, pulse=False))]),
    "src2d_tfsf": dict(size=(1.0e-6, 0.8e-6, 0.0), resolution=20e6, pml=3, courant=0.9, steps=8,
                       sources=[("TFSFSource", dict(center=(0.5e-6, 0.4e-6, 0.0), size=(0.4e-6, 0.4e-6, 0.0),
                                                    direction="-x", polarization="z", frequency=F0, pulse=False)),
                                ("TFSFSource", dict(center=(0.5e-6, 0.4e-6, 0.0), size=(0.4e-6, 0.4e-6, 0.0),
                                                    direction="+z", polarization="x", frequency=F0, pulse=False))]),
    "src2d_gauss": dict(size=(1.2e-6, 1.0e-6, 0.0), resolution=20e6, pml=3, courant=0.9, init="zero", steps=8,
                        sources=[("GaussianBeamSource", dict(center=(0.2e-6, 0.5e-6, 0.0), size=(0.0, 0.6e-6, 0.0),
                                                             direction="x", polarization="y", frequency=F0,
                                                             beam_waist=0.2e-6, pulse=True, pulse_width=3e-16))]),
    "src3d_mode": dict(size=(0.6e-6, 0.6e-6, 0.5e-6), resolution=20e6, pml=2, courant=0.5, steps=6,
                       sources=[("ModeSource", dict(center=(0.2e-6, 0.3e-6, 0.25e-6), size=(0.0, 0.4e-6, 0.3e-6),
                                                    mode=("mode", 9, 7, 1), direction="+x",
                                                    waveform=_wf("GaussianPulse", frequency=F0, pulse_width=3e-16),
                                                    amplitude=1.5, phase=0.2)),
                                ("ModeSource", dict(center=(0.3e-6, 0.3e-6, 0.2e-6), size=(0.4e-6, 0.4e-6, 0.0),
                                                    mode=("mode", 8, 8, 2), direction="-z",
                                                    waveform=_wf("ContinuousWave", frequency=F0)))]),
    # ---- monitors -----------------------------------------------------------------------------------
    "mon3d_field": dict(size=(0.6e-6, 0.5e-6, 0.4e-6), resolution=20e6, pml=2, courant=0.5, steps=8,
                        sources=[("PointSource", dict(position=(0.3e-6, 0.25e-6, 0.2e-6), component="Ey",
                                                      waveform=_wf("ContinuousWave", frequency=F0)))],
                        monitors=[("FieldMonitor", dict(center=(0.3e-6, 0.25e-6, 0.2e-6), size=(0.0, 0.3e-6, 0.2e-6),
                                                        components="all", time_domain=True,
                                                        frequencies=[0.9 * F0, F0, 1.1 * F0])),
                                  ("FieldMonitor", dict(center=(0.3e-6, 0.25e-6, 0.2e-6), size=(0.2e-6, 0.0, 0.0),
                                                        components=["Ez", "Hx"], time_domain=False,
                                                        frequencies=[F0]))]),
    "mon2d_all": dict(size=(1.0e-6, 0.8e-6, 0.0), resolution=20e6, pml=3, courant=0.9, steps=8,
                      sources=[("PlaneWaveSource", dict(center=(0.1e-6, 0.4e-6, 0.0), size=(0.0, 0.4e-6, 0.0),
                                                        direction="x", polarization="y", frequency=F0, pulse=False))],
                      monitors=[("FieldMonitor", dict(center=(0.5e-6, 0.4e-6, 0.0), size=(0.3e-6, 0.2e-6, 0.0),
                                                      components="E", time_domain=True, frequencies=[F0])),
                                ("DFTMonitor", dict(center=(0.5e-6, 0.4e-6, 0.0), size=(0.3e-6, 0.2e-6, 0.0),
                                                    frequencies=[0.95 * F0, F0, 1.05 * F0],
                                                    components=["Ex", "Ey", "Ez", "Hz"])),
                                ("FluxMonitor", dict(center=(0.5e-6, 0.4e-6, 0.0), size=(0.0, 0.4e-6, 0.0),
                                                     direction="z", frequencies=[F0, 1.1 * F0])),
                                ("FluxMonitor", dict(center=(0.5e-6, 0.4e-6, 0.0), size=(0.0, 0.4e-6, 0.0),
                                                     direction="x")),
                                ("ModeExpansionMonitor", dict(center=(0.5e-6, 0.4e-6, 0.0), size=(0.0, 0.4e-6, 0.0),
                                                              modes=[("mode", 10, 10, 3), ("mode", 7, 5, 4)],
                                                              direction="x", frequencies=[F0]))]),
    # ---- dispersive media: ADE recursions driven by E after every step (row a22) ---------------------------
    "ade3d": dict(size=(0.6e-6, 0.5e-6, 0.4e-6), resolution=20e6, pml=2, courant=0.5, steps=7,
                  sources=[("PointSource", dict(position=(0.3e-6, 0.25e-6, 0.2e-6), component="Ez",
                                                waveform=_wf("ContinuousWave", frequency=F0)))],
                  ade=[("lorentz", [(2 * 3.141592653589793 * 2.5e14, 1.0, 1e13), (2 * 3.141592653589793 * 4e14, 0.4, 3e13)], "Ez", "half"),
                       ("drude", (2 * 3.141592653589793 * 2.175e15, 2 * 3.141592653589793 * 6.5e12), "Ex", None),
                       ("debye", (2.0, 5.0, 1e-14), "Ey", "half")]),
    "ade2d": dict(size=(1.0e-6, 0.8e-6, 0.0), resolution=20e6, pml=3, courant=0.9, steps=7,
                  ade=[("lorentz", [(2 * 3.141592653589793 * 2.5e14, 1.0, 1e13)], "Ez", None),
                       ("drude", (2 * 3.141592653589793 * 2.175e15, 2 * 3.141592653589793 * 6.5e12), "Ey", "half")]),
}


# ---- BASELINE.json configs at the parity sizes of SURVEY 8(d) (too slow for the golden set: oracle-checked on the GPU box) ----
def _exact_size(n, pml, res):
    return ((n - 2 * pml) - 0.5) / res


_C1 = _exact_size(1000, 10, 50e6)
_C2 = _exact_size(121, 10, 20e6)
_C3 = (_exact_size(128, 8, 40e6), _exact_size(128, 8, 40e6), _exact_size(64, 8, 40e6))
CONFIG_SCENARIOS = {
    # config 1: 2-D 1000 x 1000 Si strip waveguide, Gaussian-beam line source, DFTMonitor (corner-patch semantics) and a
    # FieldMonitor DFT line, 11 wavelengths 1.5-1.6 um; 200 steps
    "c1": dict(size=(_C1, _C1, 0.0), resolution=50e6, pml=10, courant=0.9, init="zero", steps=200, materials="c1_strip",
               sources=[("GaussianBeamSource", dict(center=(1.2e-6, _C1 / 2, 0.0), size=(0.0, 2e-6, 0.0), direction="x",
                                                    polarization="y", frequency=193e12, beam_waist=1e-6, pulse=True,
                                                    pulse_width=10e-15))],
               monitors=[("DFTMonitor", dict(center=(16e-6, _C1 / 2, 0.0), size=(0.0, 2e-6, 0.0),
                                             frequencies=[299792458.0 / w for w in np.linspace(1.5e-6, 1.6e-6, 11)],
                                             components=["Ex", "Ey", "Hz"])),
                         ("FieldMonitor", dict(center=(16e-6, _C1 / 2, 0.0), size=(0.0, 2e-6, 0.0), components=["Ey"],
                                               time_domain=False,
                                               frequencies=[299792458.0 / w for w in np.linspace(1.5e-6, 1.6e-6, 11)]))]),
    # config 2: 3-D 121^3 (100^3 + PML 10) vacuum, TFSF +x plane wave (CW), 20 steps, white-noise start
    "c2": dict(size=(_C2, _C2, _C2), resolution=20e6, pml=10, courant=0.9, steps=20,
               sources=[("TFSFSource", dict(center=(_C2 / 2,) * 3, size=(_C2 / 2,) * 3, direction="+x", polarization="y",
                                            frequency=193e12, pulse=False))]),
    # config 3, 128 x 128 x 64 crop: Si ridge on SiO2 (heterogeneous Ca..Db), library-Si Lorentz pole in the core
    # (uncoupled recursion, like the reference), ModeSource (+x) with a .value-capable waveform; 20 steps
    "c3_crop": dict(size=_C3, resolution=40e6, pml=8, courant=0.5, steps=20, materials="c3_ridge",
                    sources=[("ModeSource", dict(center=(0.6e-6, _C3[1] / 2, _C3[2] / 2), size=(0.0, 2.0e-6, 1.0e-6),
                                                 mode=("mode", 24, 16, 5), direction="+x",
                                                 waveform=_wf("GaussianPulse", frequency=F0, pulse_width=5e-15)))],
                    monitors=[("FieldMonitor", dict(center=(2.0e-6, _C3[1] / 2, _C3[2] / 2), size=(0.0, 2.0e-6, 1.0e-6),
                                                    components=["Ey", "Hz"], time_domain=False, frequencies=[F0]))],
                    ade=[("lorentz", [(2 * 3.141592653589793 * 299792458.0 / 1.2e-6, 1.0, 1e13)], "Ez", "c3_core")]),
}


# ---- helpers -------------------------------------------------------------------------------------------
def grid_dims(spec):
    from oracle.grid import OGrid

    return OGrid(spec["size"], spec["resolution"], spec["pml"]).dims


def materials(spec):
    kind = spec.get("materials")
    if kind == "c1_strip":                     # BASELINE config 1: Si strip (eps 12.11) along x in SiO2 (2.07), 2-D
        nx, ny, _ = grid_dims(spec)
        eps = np.full((nx, ny, 1), 2.07)
        eps[:, ny // 2 - 11: ny // 2 + 11] = 12.11
        return dict(eps_rel=eps, mu_rel=np.ones_like(eps), sigma_e=np.zeros_like(eps), sigma_m=np.zeros_like(eps))
    if kind == "c3_ridge":                     # BASELINE config 3: Si ridge on an SiO2 half space, air above
        nx, ny, nz = grid_dims(spec)
        eps = np.ones((nx, ny, nz))
        eps[:, :, : nz // 2] = 2.07
        eps[:, ny // 2 - ny // 16: ny // 2 + ny // 16, nz // 2: nz // 2 + nz // 12] = 12.11
        return dict(eps_rel=eps, mu_rel=np.ones_like(eps), sigma_e=np.zeros_like(eps), sigma_m=np.zeros_like(eps))
    if kind != "random":
        return None
    dims = grid_dims(spec)
    rng = np.random.default_rng(7)
    return dict(eps_rel=1 + 11 * rng.random(dims), mu_rel=1 + 0.5 * rng.random(dims),
                sigma_e=2e3 * rng.random(dims), sigma_m=5e5 * rng.random(dims))


def initial_fields(spec, shapes):
    rng = np.random.default_rng(0)
    out = {}
    for c in COMPONENTS:
        a = rng.standard_normal(shapes[c]) * (1.0 if c[0] == "E" else 1.0 / 377.0)
        if spec.get("init") == "zero" or ("init_only" in spec and c not in spec["init_only"]):
            a = np.zeros(shapes[c])
        out[c] = a
    return out


def ade_mask(kind, shape):
    if kind is None:
        return None
    m = np.zeros(shape, dtype=bool)
    if kind == "c3_core":                      # the ridge core of materials="c3_ridge" (Ez-shaped array)
        ny, nz = shape[1] + 1, shape[2]
        m[:, ny // 2 - ny // 16: ny // 2 + ny // 16, nz // 2: nz // 2 + nz // 12] = True
        return m
    m[: shape[0] // 2] = True
    m[:, ::3] = False
    return m


def make_mode(nx, ny, seed, ns=types.SimpleNamespace):
    rng = np.random.default_rng(100 + seed)
    f = {c: rng.standard_normal((nx, ny)) + 1j * rng.standard_normal((nx, ny)) for c in COMPONENTS}
    return dict(mode_number=0, neff=2.4 + 0.01j, frequency=F0, wavelength=299792458.0 / F0,
                x=np.linspace(-0.3e-6, 0.3e-6, nx) + 0.3e-6, y=np.linspace(-0.25e-6, 0.25e-6, ny) + 0.25e-6,
                power=1.0, **f)


def _resolve(v, waveform_factory, mode_factory):
    if isinstance(v, tuple) and v and v[0] == "waveform":
        return waveform_factory(v[1], v[2])
    if isinstance(v, tuple) and v and v[0] == "mode":
        return mode_factory(make_mode(v[1], v[2], v[3]))
    if isinstance(v, list) and v and isinstance(v[0], tuple) and v[0] and v[0][0] == "mode":
        return [mode_factory(make_mode(m[1], m[2], m[3])) for m in v]
    return v


def _with_value(cls):
    """Reference ModeSource calls waveform.value(t) which no stock waveform has."""
    return type(cls.__name__ + "V", (cls,), {"value": lambda self, t: self(t)})


# ---- builders ----------------------------------------------------------------------------------------------
def build_reference(spec, prismo, backend="numpy"):
    from prismo.core.solver import FDTDSolver
    from prismo.modes.solver import WaveguideMode
    from prismo.sources import waveform as W
    import prismo.monitors.dft, prismo.monitors.flux, prismo.monitors.mode_monitor, prismo.monitors.field
    import prismo.sources.point, prismo.sources.plane_wave, prismo.sources.tfsf, prismo.sources.gaussian, prismo.sources.mode

    prismo.set_backend(backend)
    sim = prismo.Simulation(size=spec["size"], resolution=spec["resolution"], pml_layers=spec["pml"],
                            courant_factor=spec["courant"])
    m = materials(spec)
    if m is not None:
        sim.solver = FDTDSolver(sim.grid, sim.dt, m)
    classes = {"PointSource": prismo.sources.point.PointSource, "ElectricDipole": prismo.sources.point.ElectricDipole,
               "MagneticDipole": prismo.sources.point.MagneticDipole,
               "PlaneWaveSource": prismo.sources.plane_wave.PlaneWaveSource, "TFSFSource": prismo.sources.tfsf.TFSFSource,
               "GaussianBeamSource": prismo.sources.gaussian.GaussianBeamSource, "ModeSource": prismo.sources.mode.ModeSource,
               "FieldMonitor": prismo.monitors.field.FieldMonitor, "DFTMonitor": prismo.monitors.dft.DFTMonitor,
               "FluxMonitor": prismo.monitors.flux.FluxMonitor,
               "ModeExpansionMonitor": prismo.monitors.mode_monitor.ModeExpansionMonitor}
    wf = lambda kind, kw: _with_value(getattr(W, kind))(**_wf_kwargs(kw))
    mf = lambda d: WaveguideMode(**d)
    for kind, kw in spec.get("sources", []):
        sim.add_source(classes[kind](**{k: _resolve(v, wf, mf) for k, v in kw.items()}))
    for kind, kw in spec.get("monitors", []):
        sim.add_monitor(classes[kind](**{k: _resolve(v, wf, mf) for k, v in kw.items()}))
    init = initial_fields(spec, {c: sim.fields[c].shape for c in COMPONENTS})
    for c in COMPONENTS:
        sim.fields[c][...] = init[c]
    sim._ade = []
    if spec.get("ade"):
        from prismo.materials import dispersion as D
        from prismo.materials.ade import ADESolver

        for kind, params, comp, mk in spec["ade"]:
            mat = _make_material(D, kind, params)
            solver = ADESolver(mat, sim.dt, sim.fields[comp].shape)
            mask = ade_mask(mk, sim.fields[comp].shape)
            sim._ade.append((solver, comp, mask))
            if backend == "b200":
                import prismo_b200

                prismo_b200.attach_ade(sim, solver, comp, mask)
    return sim


def _make_material(D, kind, params):
    if kind == "lorentz":
        return D.LorentzMaterial(1.0, [D.LorentzPole(omega_0=w, delta_epsilon=de, gamma=g) for w, de, g in params])
    if kind == "drude":
        return D.DrudeMaterial(9.84, *params)
    return D.DebyeMaterial(*params)


def step_reference(sim, n):
    """n reference steps incl. what a user of the reference does for dispersive media: call the solver after
    every Simulation.step() (nothing in the reference calls it — SURVEY F7)."""
    for _ in range(n):
        sim.step()
        for solver, comp, mask in getattr(sim, "_ade", ()):
            if sim.solver.updater.backend.name == "b200":
                continue                                     # attached: runs on the device inside sim.step()
            e = sim.fields[comp]
            solver.update_polarization(e if mask is None else e * mask)


def build_oracle(spec):
    from oracle import sim as O, waveforms as OW

    s = O.OSimulation(spec["size"], spec["resolution"], spec["pml"], spec["courant"], materials=materials(spec))
    names = {"GaussianPulse": OW.GaussianPulse, "ContinuousWave": OW.CW, "RickerWavelet": OW.Ricker, "CustomWaveform": OW.Custom}
    wf = lambda kind, kw: names[kind](**_wf_kwargs(kw))
    mf = lambda d: types.SimpleNamespace(**d)
    for kind, kw in spec.get("sources", []):
        kw = {k: _resolve(v, wf, mf) for k, v in kw.items()}
        if kind in ("ElectricDipole", "MagneticDipole"):            # sources/point.py:76-206: a PointSource on E_p / H_p
            w = OW.make_waveform(kw["frequency"], kw.get("pulse", True), kw.get("pulse_width"),
                                 kw.get("amplitude", 1.0), kw.get("phase", 0.0))
            src = O.PointSource(kw["position"], ("E" if kind[0] == "E" else "H") + kw["polarization"], w)
        else:
            src = getattr(O, kind)(**kw)
        s.add_source(src)
    for kind, kw in spec.get("monitors", []):
        s.add_monitor(getattr(O, kind)(**{k: _resolve(v, wf, mf) for k, v in kw.items()}))
    init = initial_fields(spec, {c: s.F[c].shape for c in COMPONENTS})
    for c in COMPONENTS:
        s.F[c][...] = init[c]
    from oracle.ade import OAde

    for kind, params, comp, mk in spec.get("ade", []):
        s.ades.append(OAde(kind, params, s.dt, s.F[comp].shape, comp, ade_mask(mk, s.F[comp].shape)))
    return s


def build_mirror(spec, pb, dtype=None):
    sim = pb.Simulation(size=spec["size"], resolution=spec["resolution"], pml_layers=spec["pml"],
                        courant_factor=spec["courant"], dtype=dtype)
    m = materials(spec)
    if m is not None:
        sim.set_materials(m)
    wf = lambda kind, kw: _with_value(getattr(pb, kind))(**_wf_kwargs(kw))
    mf = lambda d: types.SimpleNamespace(**d)
    for kind, kw in spec.get("sources", []):
        sim.add_source(getattr(pb, kind)(**{k: _resolve(v, wf, mf) for k, v in kw.items()}))
    for kind, kw in spec.get("monitors", []):
        sim.add_monitor(getattr(pb, kind)(**{k: _resolve(v, wf, mf) for k, v in kw.items()}))
    init = initial_fields(spec, {c: sim.fields[c].shape for c in COMPONENTS})
    for c in COMPONENTS:
        sim.fields[c][...] = init[c]
    sim._ade = []
    for kind, params, comp, mk in spec.get("ade", []):
        solver = pb.ADESolver(_make_material(pb, kind, params), sim.dt, sim.fields[comp].shape)
        sim.add_ade(solver, comp, ade_mask(mk, sim.fields[comp].shape))
        sim._ade.append((solver, comp, None))
    return sim


def _ade_results(out, solvers):
    for n, (s, _, _) in enumerate(solvers):
        if hasattr(s, "J_current"):
            out[f"ade{n}_J"] = np.array(s.J_current)
        elif isinstance(s.P_current, list):
            for i in range(len(s.P_current)):
                out[f"ade{n}_P{i}"] = np.array(s.P_current[i])
                out[f"ade{n}_Pp{i}"] = np.array(s.P_previous[i])
        else:
            out[f"ade{n}_P"] = np.array(s.P_current)


# ---- result extraction (same keys for all three) -------------------------------------------------------------
def results_reference(sim):
    out = {"F_" + c: np.array(sim.fields[c]) for c in COMPONENTS}
    out["t"] = np.array([sim.current_time, sim.step_count], dtype=np.float64)
    for n, m in enumerate(sim.monitors):
        k = type(m).__name__
        if k == "FieldMonitor":
            if m.time_domain:
                out[f"m{n}_t"] = np.array(m._time_points)
            for c in m.components:
                if m.time_domain:
                    out[f"m{n}_td_{c}"] = np.array(m._time_data[c])
                for f in m.frequencies:
                    out[f"m{n}_fd_{c}_{f:.6e}"] = np.array(m._freq_data[c][f])
        elif k == "DFTMonitor":
            for c in m.components:
                out[f"m{n}_dft_{c}"] = np.array(m._dft_data[c])
            out[f"m{n}_steps"] = np.array([m._time_steps], dtype=np.float64)
        elif k == "FluxMonitor":
            out[f"m{n}_p"] = np.array(m._power_flow_history)
            out[f"m{n}_t"] = np.array(m._time_history)
            if m.frequencies is not None:
                for c in COMPONENTS:
                    out[f"m{n}_dft_{c}"] = np.array(getattr(m, "_dft_" + c.lower()))
                out[f"m{n}_pf"] = np.array(m.get_frequency_domain_power())
        elif k == "ModeExpansionMonitor":
            out[f"m{n}_t"] = np.array(m._time_points)
            for i in range(len(m.modes)):
                out[f"m{n}_ct_{i}"] = np.array(m._mode_coeffs_time[i])
                if m.frequencies is not None:
                    out[f"m{n}_cf_{i}"] = np.array(m._mode_coeffs_freq[i])
    _ade_results(out, getattr(sim, "_ade", ()))
    return out


results_mirror = results_reference          # the mirror classes keep the reference's attribute names


def results_oracle(s):
    out = {"F_" + c: np.array(s.F[c]) for c in COMPONENTS}
    out["t"] = np.array([s.current_time, s.step_count], dtype=np.float64)
    for n, m in enumerate(s.monitors):
        k = type(m).__name__
        if k == "FieldMonitor":
            if m.time_domain:
                out[f"m{n}_t"] = np.array(m.time_points)
            for c in m.components:
                if m.time_domain:
                    out[f"m{n}_td_{c}"] = np.array(m.time_data[c])
                for f in m.frequencies:
                    out[f"m{n}_fd_{c}_{f:.6e}"] = np.array(m.freq_data[c][f])
        elif k == "DFTMonitor":
            for c in m.components:
                out[f"m{n}_dft_{c}"] = np.array(m.dft[c])
            out[f"m{n}_steps"] = np.array([m.time_steps], dtype=np.float64)
        elif k == "FluxMonitor":
            out[f"m{n}_p"] = np.array(m.power)
            out[f"m{n}_t"] = np.array(m.times)
            if m.frequencies is not None:
                for c in COMPONENTS:
                    out[f"m{n}_dft_{c}"] = np.array(m.dft[c])
                out[f"m{n}_pf"] = np.array(m.frequency_power())
        elif k == "ModeExpansionMonitor":
            out[f"m{n}_t"] = np.array(m.times)
            for i in range(len(m.modes)):
                out[f"m{n}_ct_{i}"] = np.array(m.coeffs_time[i])
                if m.frequencies is not None:
                    out[f"m{n}_cf_{i}"] = np.array(m.coeffs_freq[i])
    for n, a in enumerate(s.ades):
        if a.kind == "drude":
            out[f"ade{n}_J"] = np.array(a.J)
        elif a.kind == "lorentz":
            for i in range(len(a.P)):
                out[f"ade{n}_P{i}"] = np.array(a.P[i])
                out[f"ade{n}_Pp{i}"] = np.array(a.Pp[i])
        else:
            out[f"ade{n}_P"] = np.array(a.P)
    return out


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return np.inf
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return float(d / n) if n > 0 else float(d)
