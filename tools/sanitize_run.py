"""Small ragged-grid runs of every sweep kernel, meant to be executed under compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Covers: two-step sweep (with mid-step sources / monitors, odd last step -> one-step sweep), heterogeneous fused sweep,
two-pass kernels, 2-D kernels, fused physics sweep with CPML, ADE ops, flux ops.  Prints SANITIZE_RUN OK at the end; the
sanitizer's own summary line says whether it saw errors."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import prismo_b200 as pb  # noqa: E402
from prismo_b200 import _lib, cpml  # noqa: E402
from prismo_b200.engine import AdeOp  # noqa: E402

C0 = 299792458.0
COMPS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


def seed(eng, rng):
    for c in COMPS:
        eng.upload(c, rng.standard_normal(eng.field_shape(c)) * (1.0 if c[0] == "E" else 1 / 377.0))


def main():
    rng = np.random.default_rng(0)
    d = 2e-8
    dt = 0.5 * d / (C0 * np.sqrt(3))
    which = sys.argv[1:] or ["tb2", "het", "twopass", "2d", "yee"]
    for dtype in ("float32", "float64"):
        dims = (21, 29, 70)
        if "tb2" in which:
            eng = pb.Engine(3, dims, (d,) * 3, dt, dtype=dtype)
            seed(eng, rng)
            sy, sz = eng.field_shape("Ey")[1:], eng.field_shape("Hz")[1:]
            eng.add_source_op(pb.SourceOp("Ey", (5, 0, 0), (6,) + sy, 0))
            eng.add_source_op(pb.SourceOp("Hz", (5, 0, 0), (6,) + sz, 1))
            eng.add_monitor_op(pb.MonitorOp("Ey", (15, 0, 0), (16,) + sy, True, 2, 0))
            eng.set_tables(7, rng.standard_normal((7, 2)), np.exp(1j * rng.standard_normal((7, 2))))
            eng.run(7)                                   # 3 pairs (two-step sweep) + 1 single step (one-step sweep)
            eng.sync()
            eng.close()
        if "het" in which:
            eng = pb.Engine(3, dims, (d,) * 3, dt, dtype=dtype)
            full = dims
            eng.set_coeffs(1 - 0.1 * rng.random(full), (dt / 8.854e-12) / (1 + 11 * rng.random(full)), 1 - 0.1 * rng.random(full),
                           np.full(full, dt / (4e-7 * np.pi)))
            seed(eng, rng)
            eng.add_ade_op(AdeOp("Ez", 0, (2, 2, 2), (9, 9, 30), 0.1, 0.1, 1.5, -0.6))
            eng.run(3)
            eng.sync()
            eng.close()
        if "twopass" in which:
            eng = pb.Engine(3, dims, (d,) * 3, dt, dtype=dtype, flags=_lib.FLAG_TWO_PASS)
            seed(eng, rng)
            eng.run(2)
            eng.sync()
            eng.close()
        if "2d" in which:
            eng = pb.Engine(2, (37, 45, 1), (d, d, 0.0), dt, dtype=dtype)
            seed(eng, rng)
            eng.run(3)
            eng.sync()
            eng.close()
        if "yee" in which:
            for fused in (0, 1, 2):
                eng = pb.Engine(3, dims, (d,) * 3, dt, dtype=dtype, flags=_lib.FLAG_YEE)
                eng.set_cpml(3, cpml.coefficient_table(dims, (d,) * 3, dt, cpml.PMLParams(thickness=3)))
                eng.set_option("yee_fused", fused)
                seed(eng, rng)
                eng.run(3)
                eng.sync()
                eng.close()
    print("SANITIZE_RUN OK", which)


if __name__ == "__main__":
    main()
