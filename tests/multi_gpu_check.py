"""Run under torchrun on N GPUs (or N processes sharing one GPU with --same-device): the peer-to-peer slab
path must reproduce a single-engine run bit for bit.  Prints 'MULTI_GPU_CHECK OK' on rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import prismo_b200 as pb
    from prismo_b200.multigpu import PeerSlabRunner, SlabStepper, balanced_slab_ranges, slab_range

    same = "--same-device" in sys.argv
    if "--sim" in sys.argv:
        return sim_check(same)
    mode = "nccl" if "--nccl" in sys.argv else "p2p"
    dtype = "float32" if "--f32" in sys.argv else "float64"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if same else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo" if same else "nccl")
    dims, steps = (8 * world + 5, 45, 70), 9
    if "--balanced" in sys.argv:                 # uneven slabs (load-balanced decomposition): same results required
        w = np.ones(dims[0])
        w[dims[0] - 3] = 9.0
        w[3] = 5.0
        spans = balanced_slab_ranges(w, world)
        slab_range = lambda nx, r, wd: spans[r]  # noqa: E731
    spacing = (2e-8, 2.5e-8, 3e-8)
    dt = 0.5 / (299792458.0 * np.sqrt(sum((1 / s) ** 2 for s in spacing)))
    comps = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
    rng = np.random.default_rng(4)
    whole_shape = {}
    for c in comps:
        n = list(dims)
        for ax in pb.grid.SHORT_AXES[c]:
            n[ax] -= 1
        whole_shape[c] = tuple(n)
    init = {c: rng.standard_normal(whole_shape[c]) * (1.0 if c[0] == "E" else 1 / 377.0) for c in comps}
    amp = np.sin(np.arange(1, steps + 1)[:, None] * np.array([[0.3, 0.7]]))
    ph = np.exp(-1j * np.arange(1, steps + 1)[:, None] * np.array([[0.2, 0.5]]))
    # first plane of the second slab: a local op for that rank AND a ghost op for the rank on its left
    src_off = int(sys.argv[sys.argv.index("--src-offset") + 1]) if "--src-offset" in sys.argv else 0
    src_plane, mon_plane = (slab_range(dims[0], 1, world)[0] + src_off) if world > 1 else dims[0] // 2, dims[0] - 3

    def ops(x0, nxl, shape_of):
        s, m = [], []
        if x0 <= src_plane < x0 + nxl:
            i = src_plane - x0
            s = [pb.SourceOp("Ey", (i, 0, 0), (i + 1,) + shape_of("Ey")[1:], 0), pb.SourceOp("Hz", (i, 0, 0), (i + 1,) + shape_of("Hz")[1:], 1)]
        elif x0 + nxl <= src_plane < x0 + nxl + 3:          # ghost ops for the two-step sweep
            i = src_plane - x0
            s = [pb.SourceOp("Ey", (i, 0, 0), (i + 1,) + shape_of("Ey")[1:], 0, ghost=True),
                 pb.SourceOp("Hz", (i, 0, 0), (i + 1,) + shape_of("Hz")[1:], 1, ghost=True)]
        if x0 <= mon_plane < x0 + nxl:
            i = mon_plane - x0
            m = [pb.MonitorOp("Ez", (i, 0, 0), (i + 1,) + shape_of("Ez")[1:], False, 2, 0)]
        return s, m

    het = "--het" in sys.argv                    # heterogeneous media + a Drude recursion whose box crosses every cut
    crng = np.random.default_rng(8)
    coef = [1 - 0.1 * crng.random(dims), (dt / 8.854187817e-12) / (1 + 11 * crng.random(dims)), 1 - 0.1 * crng.random(dims),
            np.full(dims, dt / (4e-7 * np.pi)) * (1 + 0.2 * crng.random(dims))]
    ade_box = ((1, 2, 3), (dims[0] - 2, dims[1] - 3, dims[2] - 4))

    # --aniso: coefficients painted ON THE DEVICE from a shape list (slab-local x coordinates + the neighbour's first
    # plane) with an anisotropic slab crossing every cut: per-component Cb in the het sweep, ghost plane of six arrays
    aniso = "--aniso" in sys.argv
    het = het or aniso
    gx, gy, gz = (np.arange(n) * h for n, h in zip(dims, spacing))

    def shapes():
        from prismo_b200 import geometry as G

        L = [n * h for n, h in zip(dims, spacing)]
        return [G.Box(G.Material("clad", (2.2, 2.31, 2.4)), (L[0] / 2, L[1] / 2, L[2] / 2), (2 * L[0], L[1] / 2, L[2] / 2)),
                G.Cylinder(G.Material("core", 12.1), (L[0] / 2, L[1] / 2, L[2] / 2), L[1] / 8, 0.8 * L[0], "x"),
                G.Sphere(G.Material("pad", 4.0, 1.3, 2e4, 1e3), (L[0] / 3, L[1] / 3, L[2] / 3), L[1] / 5)]

    def medium(e, x0, nxl):
        from prismo_b200.engine import AdeOp

        if not het:
            return None
        hi = min(x0 + nxl + 1, dims[0])
        if aniso:
            # --indexed: the slabs read material indices, the whole engine the coefficient arrays; else arrays everywhere
            e.set_option("het_indexed", int("--indexed" in sys.argv and nxl < dims[0]))
            e.rasterize(shapes(), gx[x0:hi], gy, gz, (1.44, 1.0, 0.0, 0.0))
            assert not np.array_equal(e.download_coeffs("Cby", hi - x0), e.download_coeffs("Cbz", hi - x0))
        else:
            e.set_coeffs(*[a[x0:hi] for a in coef])
        a, b = max(ade_box[0][0], x0), min(ade_box[1][0], x0 + e.field_shape("Ex")[0])
        if b <= a:
            return None
        return e.add_ade_op(AdeOp("Ex", 1, (a - x0,) + ade_box[0][1:], (b - x0,) + ade_box[1][1:], 0.3, 0.9)), a, b

    x0, nxl = slab_range(dims[0], rank, world)
    eng = pb.Engine(3, (nxl, dims[1], dims[2]), spacing, dt, dtype=dtype, device=local, nx_global=dims[0], x_offset=x0)
    my_ade = medium(eng, x0, nxl)
    for c in comps:
        eng.upload(c, init[c][x0:x0 + eng.field_shape(c)[0]])
    s_ops, m_ops = ops(x0, nxl, eng.field_shape)
    for o in s_ops:
        eng.add_source_op(o)
    ids = [eng.add_monitor_op(o) for o in m_ops]
    eng.set_tables(steps, amp, ph)
    runner = PeerSlabRunner(eng, rank, world) if mode == "p2p" else SlabStepper(eng, rank, world, tail_planes=3)
    runner.run(4)
    runner.run(steps - 4)
    runner.synchronize()
    mine = {c: eng.download(c) for c in comps}
    mine["dft"] = eng.dft(ids[0]) if ids else None
    mine["x0"] = x0
    mine["ade"] = (my_ade[1], my_ade[2], eng.ade_state(my_ade[0], 0)) if my_ade else None
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        whole = pb.Engine(3, dims, spacing, dt, dtype=dtype, device=local)
        whole_ade = medium(whole, 0, dims[0])
        for c in comps:
            whole.upload(c, init[c])
        s_ops, m_ops = ops(0, dims[0], whole.field_shape)
        for o in s_ops:
            whole.add_source_op(o)
        wid = [whole.add_monitor_op(o) for o in m_ops]
        whole.set_tables(steps, amp, ph)
        whole.run(steps)
        for g in gathered:
            for c in comps:
                want = whole.download(c)[g["x0"]:g["x0"] + g[c].shape[0]]
                if not np.array_equal(g[c], want):
                    ok = False
                    print(f"MISMATCH rank x0={g['x0']} {c}: max abs diff {np.abs(g[c] - want).max():.3e}")
            if g["dft"] is not None and not np.array_equal(g["dft"], whole.dft(wid[0])):
                ok = False
                print("MISMATCH dft")
            if g["ade"] is not None:
                a, b, st = g["ade"]
                want = whole.ade_state(whole_ade[0], 0)[a - ade_box[0][0]:b - ade_box[0][0]]
                if not (np.abs(st).max() > 0 and np.array_equal(st, want)):
                    ok = False
                    print(f"MISMATCH ade state planes [{a},{b})")
        print(f"MULTI_GPU_CHECK {'OK' if ok else 'FAILED'} world={world} mode={mode} dtype={dtype} dims={dims} het={het}", flush=True)
        whole.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def sim_check(same):
    """Whole Simulations (every 3-D vacuum scenario: sources, monitors) through prismo_b200.Simulation under an
    initialised process group: Session slab-decomposes them; every rank must reproduce the reference goldens."""
    import torch
    import torch.distributed as dist

    import prismo_b200 as pb
    from tests import scenarios as S

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if same else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo" if same else "nccl")
    pb.configure(device=local)
    ok = True
    names = ["upd3d_vac", "src3d_point", "src3d_plane", "src3d_tfsf", "src3d_mode", "mon3d_field", "upd3d_het", "ade3d"]
    for name in names:
        spec = S.SCENARIOS[name]
        gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        sim = S.build_mirror(spec, pb, dtype="float64")
        sim._device = local
        sim.run_steps(3)
        sim.run_steps(spec["steps"] - 3)
        assert sim.solver.updater.session().distributed == (sim.grid.dimensions[0] >= 4 * world)
        res = S.results_mirror(sim)
        for k in gold.files:
            tol = 1e-14 if name == "src3d_mode" else 0.0
            err = S.rel_l2(res[k], gold[k])
            if not (err <= tol):
                ok = False
                print(f"MISMATCH rank {rank} {name}:{k} rel-L2 {err:.3e}", flush=True)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print(f"MULTI_GPU_SIM_CHECK {'OK' if all(flags) else 'FAILED'} world={world} scenarios={len(names)}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
