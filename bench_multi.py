"""Multi-GPU leg of bench.py: the same workload slab-decomposed along x over WORLD_SIZE ranks
(one process per GPU, NCCL send/recv of 7 halo planes per interface per step, overlapped with the sweep).
Timing: CUDA events on this rank's compute stream, barrier + synchronize on both sides, MAX over ranks."""
from __future__ import annotations

import json
import os
import time

import numpy as np


def run_multi(args, rank, world, local):
    import torch
    import torch.distributed as dist

    import bench as B
    from prismo_b200.multigpu import PeerSlabRunner, SlabStepper

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name, dims = B.parse_workload(args.workload)
    cells = dims[0] * dims[1] * dims[2]
    span = None
    balanced = os.environ.get("FDTD_B200_BALANCE", "1") == "1" and not args.no_ops and world > 1
    if balanced:
        # load-balanced slabs: ranks run in lock step, so the rank that owns the DFT plane (complex128 read-modify-
        # write of 10 planes per step on top of its sweep) gets fewer planes.  Same cut on every rank (pure function).
        from prismo_b200.multigpu import balanced_slab_ranges, plane_costs

        try:
            dt0, spacing0 = B.workload_timestep()
            g_src, g_mon = B.workload_ops(dims, dt0, spacing0)
            span = balanced_slab_ranges(plane_costs(dims[0], dims[1] * dims[2], g_src, g_mon), world)[rank]
        except ValueError:                       # grid too small for >= 4 planes per rank: equal slabs
            span, balanced = None, False
    eng, dt, spacing, x0, nxl = B.make_engine(dims, args.dtype, rank, world, device=local, span=span)
    wl = B.install_workload(eng, name, dims, dt, spacing, args, x0, nxl)
    mon, mon_ids = wl["mon"], wl["mon_ids"]
    het = wl["medium"] is not None
    if het:
        eng.set_option("tb2", 0)                 # heterogeneous media: one sweep per step on every rank
    total = args.warmup + args.steps
    amp, ph, _ = B.tables(total, dt)
    eng.set_tables(total, amp, ph)
    B.seed_fields(eng, dims, x0)
    eng.sync()
    halo = os.environ.get("FDTD_B200_HALO", "p2p")
    if halo == "nccl":
        stepper = SlabStepper(eng, rank, world, tail_planes=int(os.environ.get("FDTD_B200_TAIL", "32")))
    else:
        stepper = PeerSlabRunner(eng, rank, world)

    def fence():
        stepper.synchronize()
        torch.cuda.synchronize()
        dist.barrier()

    fence()
    stepper.run(args.warmup)
    fence()
    l0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with B.ClockSampler(local) as clk:
        if halo == "nccl":
            ev0.record(stepper.compute)
            stepper.run(args.steps)
            ev1.record(stepper.compute)
            fence()
            my_ms = ev0.elapsed_time(ev1)
        else:                                   # the engine's own stream: events through the C ABI
            eng.timer_start()
            stepper.run(args.steps)
            my_ms = eng.timer_stop()
            fence()
    ms = torch.tensor([my_ms], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = torch.tensor([eng.kernel_launches - l0], device="cuda", dtype=torch.int64)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)

    timed_cs = gather_checksums(eng, rank, world, dist)
    s_params = None
    if wl["ports"]:
        def merge(mine):
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            out = {}
            for p in parts:
                out.update(p)
            return out

        s_params = B.port_s_parameters(eng, wl["ports"], wl["medium"], dims, spacing, gather=merge)

    # e2e: host buffers in / out on every rank, wall clock, max over ranks
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e_multi(eng, stepper, dims, args, mon_ids, dt, fence, dist, torch)

    check = None
    if not args.no_check:
        check = self_check_multi(eng, stepper, dims, dt, spacing, args, mon, mon_ids, x0, nxl, rank, world, fence, dist,
                                 wl["coef_fn"], wl["src_profile"])
        if rank == 0:
            import bench_check as BC

            check["timed_fields_sha"] = BC.sha_of_checksums(timed_cs)

    if rank == 0:
        n_coef = (6 if wl["aniso"] else 4) * int(het)                        # + Ca,Cb,Da,Db (+ Cb_y, Cb_z) reads
        bpc = B.BYTES_PER_CELL[args.dtype] + (4 if args.dtype == "float32" else 8) * n_coef
        peak, peak_src = B.peaks()
        value = cells * args.steps / (ms * 1e-3)
        achieved = bpc * cells * args.steps / (ms * 1e-3) / 1e9 / world
        halo_bytes = 7 * dims[1] * dims[2] * (4 if args.dtype == "float32" else 8)
        line = {"metric": "fdtd_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64",
                "data": "synthetic",
                "config": {"workload": B.workload_text(name, dims) if wl["real"] else
                                       f"{name}: 3-D {dims[0]}x{dims[1]}x{dims[2]} vacuum (uniform coefficients), TFSF +x "
                                       f"plane source, FieldMonitor DFT plane (Ey,Hz x 5 freq)",
                           "l2": "per-rank working set >> 126 MB L2 (no flush needed)",
                           "parallelism": f"x-slabs over {world} GPUs, "
                                          + (f"load-balanced (this rank: {nxl} planes), " if balanced else f"{dims[0] // world} planes each, ")
                                          + f"{halo_bytes / 1e6:.1f} MB halo per interface per step, "
                                          + ("NCCL send/recv" if halo == "nccl" else
                                             "DMA push into the neighbour's ghost planes over NVLink (CUDA IPC) + "
                                             "release/acquire flags, in-kernel wait"),
                           "kernel_path": f"heterogeneous one-step fused sweep (6 + {n_coef} arrays in, 6 out), ping-pong" if het else
                                          ("temporally blocked fused sweep (2 steps per HBM pass), ping-pong"
                                           if os.environ.get("FDTD_B200_TB2", "1") != "0" else "fused single sweep, ping-pong")},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": peak_src,
                             "kernel": ("k_fused3d_het (one step per launch)" if het else
                                        "k_fused3d_tb2 (two steps per launch)" if os.environ.get("FDTD_B200_TB2", "1") != "0"
                                        else "k_fused3d (one step per launch)")
                                       + ", per GPU, whole step incl. in-kernel halo wait (max over ranks)",
                             "note": "achieved = 48 B (fp32) per cell-update x global cells / N / step time: the per-GPU "
                                     "share of the single-GPU line's figure; traffic is not re-measured per rank (ncu is "
                                     "single-process: see the 1-GPU line / profiles/)"},
                "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches.item()), "clocks": clk.summary(), "check": check}
        if s_params is not None:
            line["s_params"] = s_params
        if wl["setup"] is not None:
            line["setup"] = wl["setup"]
        print(json.dumps(line), flush=True)
    eng.close()
    dist.destroy_process_group()


def gather_checksums(eng, rank, world, dist):
    """Per-plane checksums of every component, concatenated over the slabs in rank order (rank 0; None elsewhere)."""
    import bench as B

    mine = {c: eng.plane_checksums(c) for c in B.COMPONENTS}
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    return {c: np.concatenate([p[c] for p in parts], axis=0) for c in B.COMPONENTS}


def self_check_multi(eng, stepper, dims, dt, spacing, args, mon, mon_ids, x0, nxl, rank, world, fence, dist,
                     coef_fn=None, src_profile=None):
    """bench_check.run_check over the slab decomposition: boxes are assembled from the ranks that own their planes."""
    import bench as B
    import bench_check as BC

    def reseed(n, amp, ph):
        planes = BC.seed_fields(eng, dims, x0)
        for i in mon_ids:
            op = eng._mon_ops[i]
            eng.set_dft(i, np.zeros((op.n_freq,) + op.shape, dtype=np.complex128))
        eng.set_tables(n, amp, ph)
        eng.sync()
        dist.barrier()             # every rank's upload is complete before any rank pushes a halo (fdtd_b200.h)
        return planes

    def run_steps(n):
        stepper.run(n)
        fence()

    def fetch_box(c, lo, hi):
        a, b = max(lo[0], x0), min(hi[0], x0 + eng.field_shape(c)[0])
        piece = eng.download_box(c, (a - x0, lo[1], lo[2]), (b - x0, hi[1], hi[2])) if b > a else None
        parts = [None] * world if rank == 0 else None
        dist.gather_object(piece, parts, dst=0)
        if rank != 0:
            return None
        return np.concatenate([p for p in parts if p is not None], axis=0)

    def fetch_dft(c, lo2, hi2):
        piece = None
        names = [op.component for op in mon]
        if c in names:
            piece = eng.dft(mon_ids[names.index(c)])[:, 0, lo2[0]:hi2[0], lo2[1]:hi2[1]]
        parts = [None] * world if rank == 0 else None
        dist.gather_object(piece, parts, dst=0)
        if rank != 0:
            return None
        got = [p for p in parts if p is not None]
        return got[0] if got else None

    return BC.run_check(dims, dt, spacing, args.dtype, B.tables, dims[0] // 4, (3 * dims[0]) // 4, reseed, run_steps,
                        fetch_box, fetch_dft, lambda: gather_checksums(eng, rank, world, dist),
                        do_oracle=not args.no_ops, coef_fn=coef_fn, src_profile=src_profile)


def run_e2e_multi(eng, stepper, dims, args, mon_ids, dt, fence, dist, torch):
    import bench as B

    tdt = torch.float32 if args.dtype == "float32" else torch.float64
    host = {}
    for c in B.COMPONENTS:
        t = torch.zeros(eng.field_shape(c), dtype=tdt, pin_memory=True)
        host[c] = t.numpy()
        eng.download(c, host[c])
    amp, ph, _ = B.tables(args.steps, dt)
    fence()
    t0 = time.perf_counter()
    for c in B.COMPONENTS:
        eng.upload(c, host[c])
    eng.set_tables(args.steps, amp, ph)
    eng.sync()
    dist.barrier()                 # every rank's upload is complete before any rank pushes a halo (fdtd_b200.h)
    stepper.run(args.steps)
    stepper.synchronize()
    d2h = 0
    for c in B.COMPONENTS:
        eng.download(c, host[c])
        d2h += host[c].nbytes
    for i in mon_ids:
        d2h += eng.dft(i).nbytes
    el = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    bytes_ = torch.tensor([sum(a.nbytes for a in host.values()) + amp.nbytes + ph.nbytes, d2h], device="cuda", dtype=torch.float64)
    dist.all_reduce(bytes_, op=dist.ReduceOp.SUM)
    el = float(el.item())
    cells = dims[0] * dims[1] * dims[2]
    return {"value": cells * args.steps / el, "unit": "cell-updates/s", "h2d_bytes_per_step": float(bytes_[0].item()) / args.steps,
            "d2h_bytes_per_step": float(bytes_[1].item()) / args.steps, "seconds": el,
            "what": "every rank: upload its slab of 6 fields from pinned host memory + tables, K steps with halo "
                    "exchange, download fields + DFT planes; wall clock, max over ranks"}
