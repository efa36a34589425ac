"""GPU experiment (not part of pytest): the opt-in fused physics sweep (fdtd_yee_fused.cuh, option "yee_fused")
against the default two-pass physics kernels — bitwise in fp64, rel-L2 in fp32 — and their throughput.

    gpurun -- python tools/check_yee_fused.py [N]         (N = cube edge for the timing part, default 512)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import prismo_b200 as pb  # noqa: E402
from prismo_b200 import _lib, cpml  # noqa: E402

C0 = 299792458.0
COMPS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")


def run(dims, dtype, thickness, fused, steps, seed=1, lx=None):
    d = 2e-8
    dt = 0.9 * d / (C0 * np.sqrt(3))
    eng = pb.Engine(3, dims, (d,) * 3, dt, dtype=dtype, flags=_lib.FLAG_YEE)
    if thickness:
        eng.set_cpml(thickness, cpml.coefficient_table(dims, (d,) * 3, dt, cpml.PMLParams(thickness=thickness, alpha_max=0.05)))
    eng.set_option("yee_fused", int(fused))
    if lx:
        eng.set_option("fused_lx", lx)
    rng = np.random.default_rng(seed)
    for c in COMPS:
        shp = eng.field_shape(c)
        eng.upload(c, rng.standard_normal(shp) * (1.0 if c[0] == "E" else 1 / 377.0))
    eng.run(steps)
    out = {c: eng.download(c) for c in COMPS}
    eng.close()
    return out


def main():
    ok = True
    for dims in ((28, 24, 26), (37, 33, 70), (9, 8, 7), (64, 47, 130)):
        for thickness in (0, 3):
            if 2 * thickness + 1 > min(dims):
                continue
            for steps in (1, 7, 40):
                for lx in (None, 5):
                    a = run(dims, "float64", thickness, False, steps)
                    b = run(dims, "float64", thickness, True, steps, lx=lx)
                    bad = [c for c in COMPS if not np.array_equal(a[c], b[c])]
                    if bad:
                        ok = False
                        worst = max(np.abs(a[c] - b[c]).max() / (np.abs(a[c]).max() + 1e-300) for c in bad)
                        print(f"MISMATCH fp64 dims={dims} t={thickness} steps={steps} lx={lx}: {bad} worst rel {worst:.3e}")
            a = run(dims, "float64", thickness, False, 20)
            b = run(dims, "float32", thickness, True, 20)
            err = max(np.linalg.norm(a[c] - b[c]) / np.linalg.norm(a[c]) for c in COMPS)
            print(f"fp32 fused vs fp64 two-pass dims={dims} t={thickness}: rel-L2 {err:.2e}")
            ok &= err < 1e-4
    print("YEE_FUSED_CHECK", "OK" if ok else "FAILED")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    d = 2e-8
    dt = 0.9 * d / (C0 * np.sqrt(3))
    for fused in (0, 1):
        eng = pb.Engine(3, (n, n, n), (d,) * 3, dt, dtype="float32", flags=_lib.FLAG_YEE)
        eng.set_cpml(10, cpml.coefficient_table((n, n, n), (d,) * 3, dt, cpml.PMLParams(thickness=10)))
        eng.set_option("yee_fused", fused)
        eng.run(6)
        eng.sync()
        eng.timer_start()
        eng.run(40)
        ms = eng.timer_stop()
        g = n ** 3 * 40 / ms / 1e6
        print(f"physics mode {n}^3 fp32 {'fused' if fused else 'two-pass'}: {g:.1f} Gcell/s = {g * 48 / 6547.2:.3f} of the 48-B roofline")
        eng.close()


if __name__ == "__main__":
    main()
