"""fdtd_b200 -- B200-native engine for the per-timestep Yee update of flaport/fdtd.

Drop-in for the reference's hot path and nothing else:

    import fdtd_b200 as fdtd
    fdtd.set_backend("cuda.float32")          # or "cuda" / "cuda.float64"
    grid = fdtd.Grid((1024, 1024, 1024), grid_spacing=77.5e-9)
    grid[0:10, :, :] = fdtd.PML(); ...
    grid[512, 512, 512] = fdtd.PointSource(period=20)
    grid.run(200)

`Grid(...)`, `grid[...] = PML / PeriodicBoundary / Object / AbsorbingObject / AnisotropicObject /
PointSource / LineSource / PlaneSource / SoftArbitraryPointSource / LineDetector / BlockDetector /
CurrentDetector`, `FrequencyRoutines`, `grid.run() / step() /
update_E() / update_H() / reset()`, `grid.E / grid.H` and the detector outputs behave as in the
reference (fdtd/__init__.py:6-14); each half-step is one fused sm_100a CUDA kernel reached
through the C ABI of include/fdtd_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from .backend import backend, set_backend
from .grid import Grid
from .sources import PointSource, LineSource, PlaneSource, SoftArbitraryPointSource
from .detectors import LineDetector, BlockDetector, CurrentDetector
from .objects import Object, AbsorbingObject, AnisotropicObject
from .boundaries import PeriodicBoundary, PML, DomainBorderPML
from .fourier import FrequencyRoutines
from .visualization import dB_map_2D, plot_detection
from . import constants, conversions, waveforms
