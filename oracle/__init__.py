"""CPU oracle for the FDTD time-step path of rithulkamesh/prismo.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference's algorithm for the one hot path this
repository accelerates (``Simulation.step`` -> ``MaxwellUpdater.step`` plus the sources and
monitors that run inside the step).  Every function cites the reference file:line it follows.

Rules ("Oracle"):
  * Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import anything from here.  ``prismo_b200`` never does.
  * Parity is PINNED: ``tests/test_oracle_vs_reference.py`` runs this restatement against the
    real reference (imported from /root/reference through ``oracle/_shim``) on random inputs,
    and ``tests/golden/*.npz`` hold outputs of the real reference (``oracle/make_golden.py``)
    that travel to the GPU box where /root/reference does not exist.
"""
from . import kernels, grid, waveforms, sim, ade  # noqa: F401
