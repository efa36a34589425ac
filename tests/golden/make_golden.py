"""Generate tests/golden/*.npz by running the REAL reference (rithulkamesh/prismo, unmodified).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Each file holds every result of one scenario in tests/scenarios.py (final fields, counters, every
monitor output) after ``steps`` calls of the reference's own ``Simulation.step()`` on its NumPy backend.
The fixtures travel to the GPU box, where the reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests import scenarios as S  # noqa: E402


def main():
    prismo = ref_loader.load()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, spec in S.SCENARIOS.items():
        sim = S.build_reference(spec, prismo)
        S.step_reference(sim, spec["steps"])
        res = S.results_reference(sim)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **res)
        print(f"{name}: grid {sim.grid.dimensions}, {len(res)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


def main_tensor():
    """Row a23: AnisotropicUpdater.update_e_from_curl_h / update_h_from_curl_e of the real reference on the cases of
    tests/tensor_cases.py, float64 and float32 inputs -> tests/golden/aniso.npz (outputs only; inputs are seeded)."""
    ref_loader.load()
    from prismo.materials.tensor import AnisotropicUpdater, TensorComponents, TensorMaterial

    from tests import tensor_cases as T

    res = {}
    for name, spec in T.cases().items():
        mat = TensorMaterial(TensorComponents(**spec["eps"]), None if spec["mu"] is None else TensorComponents(**spec["mu"]))
        upd = AnisotropicUpdater(mat, T.DT)
        for dt in ("float64", "float32"):
            f, c = T.inputs(dt)
            for which, fn in (("e", upd.update_e_from_curl_h), ("h", upd.update_h_from_curl_e)):
                for k, a in enumerate(fn(tuple(f), tuple(c))):
                    res[f"{name}.{dt}.{which}{k}"] = np.asarray(a)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "aniso.npz")
    np.savez_compressed(path, **res)
    print(f"aniso: {len(res)} arrays, dtypes {sorted({str(a.dtype) for a in res.values()})}, {os.path.getsize(path) / 1024:.0f} KiB")


def main_raster():
    """Row f4: Shape.rasterize / GeometryGroup.rasterize masks and MaxwellUpdater's Ca..Db of the real reference on the
    isotropic scenes of tests/raster_cases.py -> tests/golden/raster.npz."""
    prismo = ref_loader.load()
    from tests import raster_cases as R

    res = {}
    for name in ("scene3d", "scene2d"):
        masks, coefs = R.reference_scene(name, prismo)
        res[f"{name}.masks"] = np.packbits(np.stack(masks).astype(np.uint8), axis=None)
        res[f"{name}.n_masks"] = np.array(len(masks))
        for k, a in zip(("Ca", "Cb", "Da", "Db"), coefs):
            res[f"{name}.{k}"] = a
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raster.npz")
    np.savez_compressed(path, **res)
    print(f"raster: {len(res)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    if "--raster-only" in sys.argv:
        main_raster()
        sys.exit(0)
    if "--tensor-only" not in sys.argv:
        main()
    main_tensor()
    main_raster()
