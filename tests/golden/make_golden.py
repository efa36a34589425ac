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


if __name__ == "__main__":
    main()
