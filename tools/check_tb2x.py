"""GPU check: the TMA / mbarrier two-step sweep (fdtd_tb2x.cuh) against the register-prefetch one (fdtd_tb2.cuh) and the
one-step sweep, bitwise, on ragged grids, with mid-step sources / monitors.  Prints TB2X_CHECK OK / FAILED."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import prismo_b200 as pb  # noqa: E402

C0 = 299792458.0
COMPS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


def run(dims, dtype, steps, opts, ops=True, lx=0, seed=2):
    d = 2e-8
    dt = 0.5 * d / (C0 * np.sqrt(3))
    eng = pb.Engine(3, dims, (d,) * 3, dt, dtype=dtype)
    for k, v in opts.items():
        eng.set_option(k, v)
    if lx:
        eng.set_option("fused_lx", lx)
    rng = np.random.default_rng(seed)
    for c in COMPS:
        eng.upload(c, rng.standard_normal(eng.field_shape(c)) * (1.0 if c[0] == "E" else 1 / 377.0))
    mid = None
    if ops:
        sy, sz = eng.field_shape("Ey")[1:], eng.field_shape("Hz")[1:]
        p = dims[0] // 3
        eng.add_source_op(pb.SourceOp("Ey", (p, 0, 0), (p + 1,) + sy, 0))
        eng.add_source_op(pb.SourceOp("Hz", (p, 0, 0), (p + 1,) + sz, 1))
        q = (2 * dims[0]) // 3
        mid = eng.add_monitor_op(pb.MonitorOp("Ez", (q, 1, 1), (q + 1, dims[1] - 2, dims[2] - 2), True, 2, 0))
        eng.set_tables(steps, np.sin(np.arange(1, steps + 1)[:, None] * np.array([[0.3, 0.7]])),
                       np.exp(-1j * np.arange(1, steps + 1)[:, None] * np.array([[0.2, 0.5]])))
    eng.run(steps)
    out = {c: eng.download(c) for c in COMPS}
    if mid is not None:
        out["dft"] = eng.dft(mid)
        out["rec"] = eng.records(mid, steps)
    eng.close()
    return out


def main():
    ok = True
    cases = [((24, 40, 70), 0), ((40, 47, 130), 0), ((33, 16, 121), 7), ((9, 31, 250), 3), ((70, 15, 64), 1),
             ((64, 100, 300), 16), ((130, 20, 64), 0)]
    for dims, lx in cases:
        for dtype in ("float32", "float64"):
            for ops in (False, True):
                steps = 7
                ref = run(dims, dtype, steps, {"tb2": 1, "tb2x": 0}, ops, lx)
                for S, D in ((4, 4), (3, 2), (5, 3)):
                    got = run(dims, dtype, steps, {"tb2": 1, "tb2x": 1, "tb2x_stages": S, "tb2x_slots": D}, ops, lx)
                    bad = [k for k in ref if not np.array_equal(ref[k], got[k], equal_nan=True)]
                    if bad:
                        ok = False
                        k = bad[0]
                        w = np.argwhere(ref[k] != got[k])
                        print(f"MISMATCH dims={dims} lx={lx} {dtype} ops={ops} S={S} D={D}: {bad}; first {k}{tuple(w[0])} "
                              f"n={len(w)} lo={w.min(0)} hi={w.max(0)}", flush=True)
        print("done", dims, flush=True)
    print("TB2X_CHECK", "OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
