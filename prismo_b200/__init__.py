"""prismo_b200 — a B200-native FDTD time-stepping engine behind Prismo's backend API.

Two ways in (both end in libfdtd_b200.so; there is no CPU stepping path):
  * ``prismo_b200.register()`` then ``prismo.set_backend("b200")``: unchanged reference ``Simulation``,
    source, material and monitor objects run on the GPU (plugin.py);
  * the interface-compatible host classes exported here (same names and arguments as the reference),
    for environments where the reference package is not installed.
``Engine`` is the low-level object over the C ABI (include/fdtd_b200.h).
"""
from .grid import GridSpec, YeeGrid
from .waveform import ContinuousWave, CustomWaveform, GaussianPulse, RickerWavelet, Waveform
from .sources import (ElectricDipole, GaussianBeamSource, MagneticDipole, ModeSource, PlaneWaveSource, PointSource,
                      Source, TFSFSource)
from .monitors import DFTMonitor, FieldMonitor, FluxMonitor, ModeExpansionMonitor, Monitor
from .materials import (ADESolver, AnisotropicUpdater, DebyeMaterial, DrudeMaterial, LorentzMaterial, LorentzPole,
                        TensorComponents, TensorMaterial, attach_ade, tensor_update)
from .cpml import PMLParams
from .geometry import Box, Cylinder, GeometryGroup, Material, Polygon, Shape, Sphere
from .session import Session, configure
from .simulation import ElectromagneticFields, FDTDSolver, MaxwellUpdater, Simulation
from .engine import Engine, MonitorOp, SourceOp
from .plugin import clear_geometry, register, set_geometry

__version__ = "0.1.0"
from .sweep import ParameterSweep, SweepParameter

__all__ = [n for n in dir() if not n.startswith("_")]
