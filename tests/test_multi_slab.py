"""x-slab decomposition.  CPU: world_size-2 and -3 gloo runs of SlabStepper on a NumPy slab engine, compared
bitwise with the undecomposed oracle.  GPU (1 device): two real engines as two slabs of one grid with the halo
copied device-to-device, compared bitwise with a single engine — the kernel-side slab semantics."""
import os
import socket

import numpy as np
import pytest

from tests import scenarios as S

DIMS = (23, 12, 14)
SPACING = (2e-8, 2.5e-8, 3e-8)
DT = 0.5 / (299792458.0 * np.sqrt(sum((1 / s) ** 2 for s in SPACING)))
STEPS = 6
F0 = 193.4e12


def _initial(dims):
    rng = np.random.default_rng(11)
    out = {}
    for c in S.COMPONENTS:
        n = list(dims)
        for ax in {"Ex": (1, 2), "Ey": (0, 2), "Ez": (0, 1), "Hx": (0,), "Hy": (1,), "Hz": (2,)}[c]:
            n[ax] -= 1
        out[c] = rng.standard_normal(n) * (1.0 if c[0] == "E" else 1 / 377.0)
    return out


def _tables(n):
    t, amp, ph = 0.0, np.zeros((n, 2)), np.zeros((n, 2), dtype=np.complex128)
    for s in range(n):
        t += DT
        amp[s] = (-np.sin(2 * np.pi * F0 * t), 0.5 * np.cos(2 * np.pi * F0 * t))
        ph[s] = np.exp(-1j * 2 * np.pi * np.array([F0, 1.1 * F0]) * t)
    return amp, ph


def _global_ops(pb):
    """A whole-plane source on global plane 7 (Ey) and 15 (Hz), a DFT plane at global plane 11 (Ez)."""
    nx, ny, nz = DIMS
    src = [("Ey", 7, (ny, nz - 1), 0), ("Hz", 15, (ny, nz - 1), 1)]
    mon = [("Ez", 11, (ny - 1, nz))]
    return src, mon


def _local_ops(pb, x0, nxl):
    src, mon = _global_ops(pb)
    s_ops = [pb.SourceOp(c, (p - x0, 0, 0), (p - x0 + 1,) + shp, tab) for c, p, shp, tab in src if x0 <= p < x0 + nxl]
    m_ops = [pb.MonitorOp(c, (p - x0, 0, 0), (p - x0 + 1,) + shp, False, 2, 0) for c, p, shp in mon if x0 <= p < x0 + nxl]
    return s_ops, m_ops


def _reference_run():
    """Undecomposed oracle run with the same ops."""
    from oracle import kernels

    F = _initial(DIMS)
    coeffs = kernels.vacuum_coefficients(DIMS, DT)
    amp, ph = _tables(STEPS)
    dft = np.zeros((2, DIMS[1] - 1, DIMS[2]), dtype=np.complex128)
    for s in range(STEPS):
        kernels.step(F, coeffs, SPACING, False)
        F["Ey"][7] += amp[s, 0]
        F["Hz"][15] += amp[s, 1]
        d = F["Ez"][11].copy()
        for k in range(2):
            dft[k] += (d * ph[s, k].real) * DT + 1j * ((d * ph[s, k].imag) * DT)
    return F, dft


def _uneven_spans(world):
    """Load-balanced slabs for a grid whose DFT plane (11) and source planes (7, 15) are expensive."""
    from prismo_b200.multigpu import balanced_slab_ranges

    w = np.ones(DIMS[0])
    w[11] += 6.0
    w[7] += 2.0
    w[15] += 2.0
    return balanced_slab_ranges(w, world, min_planes=4)


def _worker(rank, world, port, out_dir, uneven=False):
    import torch.distributed as dist

    import prismo_b200 as pb
    from prismo_b200.multigpu import SlabStepper, slab_range
    from tests._fake_engine import FakeSlabEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x0, nxl = _uneven_spans(world)[rank] if uneven else slab_range(DIMS[0], rank, world)
    eng = FakeSlabEngine(3, (nxl, DIMS[1], DIMS[2]), SPACING, DT, nx_global=DIMS[0], x_offset=x0)
    init = _initial(DIMS)
    for c in S.COMPONENTS:
        eng.upload(c, init[c][x0:x0 + eng.field_shape(c)[0]])
    s_ops, m_ops = _local_ops(pb, x0, nxl)
    for o in s_ops:
        eng.add_source_op(o)
    ids = [eng.add_monitor_op(o) for o in m_ops]
    amp, ph = _tables(STEPS)
    eng.set_tables(STEPS, amp, ph)
    st = SlabStepper(eng, rank, world, tail_planes=3)
    st.run(STEPS)
    res = {c: eng.download(c) for c in S.COMPONENTS}
    for i in ids:
        res["dft"] = eng.dft(i)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x0=x0, **res)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,uneven", [(2, False), (3, False), (3, True)])
def test_slab_stepper_gloo_bitwise(world, uneven, tmp_path):
    import torch.multiprocessing as mp

    if uneven:
        sizes = [n for _, n in _uneven_spans(world)]
        assert max(sizes) - min(sizes) >= 2                     # really uneven
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), uneven), nprocs=world, join=True)
    F, dft = _reference_run()
    seen_dft = False
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        x0 = int(z["x0"])
        for c in S.COMPONENTS:
            got = z[c]
            assert np.array_equal(got, F[c][x0:x0 + got.shape[0]]), f"rank {r} {c}"
        if "dft" in z.files:
            assert np.array_equal(z["dft"][:, 0], dft)
            seen_dft = True
    assert seen_dft


def test_slab_ranges_cover_the_grid():
    from prismo_b200.multigpu import slab_range

    for nx in (8, 23, 1024):
        for world in (1, 2, 3, 8):
            spans = [slab_range(nx, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == nx
            assert all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def test_balanced_slab_ranges():
    from prismo_b200 import MonitorOp, SourceOp
    from prismo_b200.multigpu import balanced_slab_ranges, plane_costs, slab_range

    rng = np.random.default_rng(5)
    for nx in (8, 23, 128, 1024):
        for world in (1, 2, 3, 8):
            if nx < 4 * world:
                with pytest.raises(ValueError):
                    balanced_slab_ranges(np.ones(nx), world)
                continue
            # unit costs: as even as slab_range
            spans = balanced_slab_ranges(np.ones(nx), world)
            assert sorted(n for _, n in spans) == sorted(slab_range(nx, r, world)[1] for r in range(world))
            # random costs: a partition, >= 4 planes each, never worse than the uniform cut
            cost = 1.0 + 30.0 * (rng.random(nx) < 0.02) * rng.random(nx)
            spans = balanced_slab_ranges(cost, world)
            assert spans[0][0] == 0 and sum(n for _, n in spans) == nx
            assert all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert min(n for _, n in spans) >= 4
            worst = max(cost[a:a + n].sum() for a, n in spans)
            uniform = max(cost[a:a + n].sum() for a, n in (slab_range(nx, r, world) for r in range(world)))
            assert worst <= uniform + 1e-9
    # the bench workload: the rank that owns the DFT plane gets fewer planes
    src = [SourceOp("Ey", (256, 0, 0), (257, 1024, 1023), 0), SourceOp("Hz", (256, 0, 0), (257, 1023, 1023), 1)]
    mon = [MonitorOp("Ey", (768, 0, 0), (769, 1024, 1023), False, 5, 0), MonitorOp("Hz", (768, 0, 0), (769, 1023, 1023), False, 5, 0)]
    cost = plane_costs(1024, 1024 * 1024, src, mon)
    assert cost[0] == 1.0 and 15.9 < cost[256] < 16.1 and 33.0 < cost[768] < 34.4
    spans = balanced_slab_ranges(cost, 8)
    owner = [n for a, n in spans if a <= 768 < a + n][0]
    assert owner < 110 and max(cost[a:a + n].sum() for a, n in spans) < 138


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_two_slabs_on_one_gpu_match_single_engine(dtype):
    """Kernel-side slab semantics (x_offset / nx_global masks, ghost planes, H+ recompute on the ghost plane):
    two engines on one device, halo copied device-to-device each step, bitwise equal to one engine."""
    import torch

    import prismo_b200 as pb
    from prismo_b200.multigpu import HALO_SPEC, SlabStepper, _CudaAlias, slab_range

    dims = (37, 33, 70)
    rng = np.random.default_rng(2)
    whole = pb.Engine(3, dims, SPACING, DT, dtype=dtype)
    init = {c: rng.standard_normal(whole.field_shape(c)) * (1.0 if c[0] == "E" else 1 / 377.0) for c in S.COMPONENTS}
    for c in S.COMPONENTS:
        whole.upload(c, init[c])
    whole.run(5)
    slabs = []
    for r in range(2):
        x0, nxl = slab_range(dims[0], r, 2)
        e = pb.Engine(3, (nxl, dims[1], dims[2]), SPACING, DT, dtype=dtype, nx_global=dims[0], x_offset=x0)
        for c in S.COMPONENTS:
            e.upload(c, init[c][x0:x0 + e.field_shape(c)[0]])
        slabs.append((e, x0, nxl))
    esz = np.dtype(dtype).itemsize
    ts = "<f4" if esz == 4 else "<f8"
    for _ in range(5):
        left, right = slabs[0][0], slabs[1][0]
        left.sync(); right.sync()
        for comp, planes in HALO_SPEC:
            first, _, pbytes = right.halo_ptrs(comp)
            _, ghost, _ = left.halo_ptrs(comp)
            n = planes * pbytes // esz
            torch.as_tensor(_CudaAlias(ghost, n, ts), device="cuda").copy_(torch.as_tensor(_CudaAlias(first, n, ts), device="cuda"))
        torch.cuda.synchronize()
        for e, _, nxl in slabs:
            e.sweep(0, nxl - 4, False)
            e.sweep(nxl - 4, nxl, True)
            e.post_step()
            e.sync()
    for e, x0, nxl in slabs:
        for c in S.COMPONENTS:
            got = e.download(c)
            assert np.array_equal(got, whole.download(c)[x0:x0 + got.shape[0]]), (x0, c)
        e.close()
    whole.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "sim", "het", "aniso", "indexed"])
def test_peer_to_peer_slabs_two_processes_one_gpu(mode):
    """The production halo protocol (CUDA IPC mappings, DMA push, release/acquire flags, in-kernel wait) with two
    processes time-slicing one GPU: bitwise equal to a single engine.  Halo waits time out after 5 s."""
    import subprocess
    import sys

    env = dict(os.environ, FDTD_B200_HALO_TIMEOUT_MS="5000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(os.path.dirname(__file__), "multi_gpu_check.py"),
           "--same-device"] + {"sim": ["--sim"], "het": ["--het"], "aniso": ["--aniso"], "indexed": ["--aniso", "--indexed"], "p2p": []}[mode]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    want = "MULTI_GPU_SIM_CHECK OK" if mode == "sim" else "MULTI_GPU_CHECK OK"
    assert want in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("args", [[], ["--f32"], ["--balanced"], ["--sim"], ["--src-offset", "2"], ["--het"], ["--aniso"],
                                  ["--aniso", "--indexed"]],
                         ids=["p2p", "f32", "balanced", "sim", "ghost-src", "het", "aniso", "indexed"])
def test_peer_to_peer_slabs_on_all_gpus(args):
    """The same check on REAL separate GPUs (NCCL rendezvous, CUDA-IPC peer mappings over NVLink): one rank per
    device, every device of the box.  Skips on a single-GPU box."""
    import subprocess
    import sys

    from tests.conftest import _cuda_devices

    n = _cuda_devices()
    if n < 2:
        pytest.skip("needs >= 2 CUDA devices")
    env = dict(os.environ, FDTD_B200_HALO_TIMEOUT_MS="20000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(os.path.dirname(__file__), "multi_gpu_check.py")] + args
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    want = "MULTI_GPU_SIM_CHECK OK" if "--sim" in args else "MULTI_GPU_CHECK OK"
    assert want in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---- whole Simulations (sources + monitors) slab-decomposed through Session / SlabExecutor -----------------------
DIST_SCENARIOS = ["upd3d_vac", "src3d_point", "src3d_plane", "src3d_tfsf", "src3d_mode", "mon3d_field"]


def _sim_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import prismo_b200 as pb
    import prismo_b200.session as session
    from tests._fake_engine import FakeSlabEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    session.Engine = FakeSlabEngine
    for name in DIST_SCENARIOS:
        spec = S.SCENARIOS[name]
        sim = S.build_mirror(spec, pb)
        sim.run_steps(3)
        sim.run_steps(spec["steps"] - 3)
        assert sim.solver.updater.session().distributed
        np.savez(os.path.join(out_dir, f"{name}_rank{rank}.npz"), **S.results_mirror(sim))
    dist.barrier()
    dist.destroy_process_group()


def test_simulation_run_is_slab_decomposed_under_torch_distributed(tmp_path):
    """Every rank builds the same Simulation; Session sees an initialised process group and runs its x-slab only
    (ops clipped, monitor pieces re-assembled).  Every rank must end with the single-process reference results."""
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_sim_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    for name in DIST_SCENARIOS:
        gold = np.load(os.path.join(gold_dir, name + ".npz"))
        for r in range(world):
            got = np.load(tmp_path / f"{name}_rank{r}.npz")
            assert sorted(got.files) == sorted(gold.files)
            for k in gold.files:
                if name == "src3d_mode":
                    assert S.rel_l2(got[k], gold[k]) <= 1e-14, (name, r, k)
                else:
                    assert np.array_equal(got[k], gold[k], equal_nan=True), (name, r, k, S.rel_l2(got[k], gold[k]))


# ---- device-painted geometry on slabs: every rank paints its planes + the neighbour's first one ---------------------
_GEO_SPEC = "src3d_point"


def _geo_shapes(G, aniso):
    return [G.Box(G.Material("clad", (2.2, 2.31, 2.4) if aniso else 2.2), (0.3e-6, 0.3e-6, 0.1e-6), (2e-6, 0.5e-6, 0.25e-6)),
            G.Sphere(G.Material("ball", 11.9, 1.3, 40.0, 5.0), (0.3e-6, 0.25e-6, 0.35e-6), 0.16e-6),
            G.GeometryGroup([G.Cylinder(G.Material("rod", 4.0), (0.3e-6, 0.3e-6, 0.3e-6), 0.12e-6, 2e-6, "x"),
                             G.Box(G.Material("cut", 1.0), (0.3e-6, 0.3e-6, 0.3e-6), (0.2e-6, 1e-6, 1e-6))], "difference")]


def _geo_run(pb, aniso):
    from prismo_b200 import geometry as G

    sim = S.build_mirror(S.SCENARIOS[_GEO_SPEC], pb)
    sess = sim.solver.updater.session()
    sess.set_geometry(_geo_shapes(G, aniso))
    sim.run_steps(3)
    sim.run_steps(S.SCENARIOS[_GEO_SPEC]["steps"] - 3)
    return sim, sess


def _geo_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import prismo_b200 as pb
    import prismo_b200.session as session
    from tests._fake_engine import FakeSlabEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    session.Engine = FakeSlabEngine
    for aniso in (False, True):
        sim, sess = _geo_run(pb, aniso)
        assert sess.distributed
        ex = sess.engine
        planes = ex.nxl + (1 if rank < world - 1 else 0)
        np.savez(os.path.join(out_dir, f"geo{int(aniso)}_rank{rank}.npz"), x0=ex.x0,
                 Cb=ex.download_coeffs("Cb", planes), **(dict(Cbz=ex.download_coeffs("Cbz", planes)) if aniso else {}),
                 **S.results_mirror(sim))
    dist.barrier()
    dist.destroy_process_group()


def test_set_geometry_is_slab_decomposed_under_torch_distributed(tmp_path, monkeypatch):
    """Session.set_geometry under an initialised process group: each rank rasterises its slice of the x coordinates (+1
    plane); results equal the single-process run on the whole grid, which equals the oracle on host-painted arrays."""
    import torch.multiprocessing as mp

    import prismo_b200 as pb
    import prismo_b200.session as session
    from oracle import kernels, raster
    from prismo_b200 import geometry as G
    from tests._fake_engine import FakeEngine, shape_dict

    world = 2
    mp.spawn(_geo_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    monkeypatch.setattr(session, "Engine", FakeEngine)
    for aniso in (False, True):
        sim, sess = _geo_run(pb, aniso)
        assert not sess.distributed
        want = S.results_mirror(sim)
        whole_cb = sess.engine.download_coeffs("Cb")
        # the single-process run against the oracle on host-painted arrays
        o = S.build_oracle(S.SCENARIOS[_GEO_SPEC])
        x, y, z = G.cell_coordinates(sim.grid)
        o.coeffs = raster.coefficient_arrays([shape_dict(s) for s in _geo_shapes(G, aniso)], x, y, z, sim.dt)
        assert isinstance(o.coeffs[1], tuple) == aniso
        o.run_steps(S.SCENARIOS[_GEO_SPEC]["steps"])
        oo = S.results_oracle(o)
        for k in oo:
            assert np.array_equal(want[k], oo[k]), (aniso, k)
        assert 0 < (whole_cb != whole_cb.flat[0]).sum() < whole_cb.size
        for r in range(world):
            got = np.load(tmp_path / f"geo{int(aniso)}_rank{r}.npz")
            x0 = int(got["x0"])
            assert np.array_equal(got["Cb"], whole_cb[x0:x0 + got["Cb"].shape[0]]), (aniso, r)
            if aniso:
                assert not np.array_equal(got["Cbz"], got["Cb"])
            for k in want:
                assert np.array_equal(got[k], want[k], equal_nan=True), (aniso, r, k)
