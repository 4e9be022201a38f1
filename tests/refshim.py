"""pytest plugin (TEST INFRASTRUCTURE) that lets the REFERENCE's own test-suite run against fdtd_b200:

    python -m pytest /root/reference/tests -p refshim -p no:cacheprovider      (PYTHONPATH = this directory)

`fdtd` and its submodules are aliased to fdtd_b200 (here with the CPU interpreter build of the kernels, on a GPU box
with the CUDA library), `fdtd.backend.backend_names` is what the reference's conftest parametrises over, and
`set_backend(<reference backend name>)` keeps the engine under test instead of selecting a CPU backend of the
reference.  matplotlib (absent from this image, imported at module level by two of the reference's test files) is an
empty stand-in.  Nothing from the reference is copied: its tests are collected where they lie."""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

for _m in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors"):
    sys.modules.setdefault(_m, types.ModuleType(_m))

import torch  # noqa: E402

if torch.cuda.is_available():
    import fdtd_b200
    fdtd_b200.set_backend("cuda.float64")
else:
    from emu.harness import use_emu
    fdtd_b200 = use_emu("float64")

sys.modules["fdtd"] = fdtd_b200
for _name in ("backend", "grid", "sources", "detectors", "objects", "boundaries", "fourier", "waveforms", "constants",
              "conversions", "visualization"):
    sys.modules["fdtd." + _name] = importlib.import_module("fdtd_b200." + _name)

_backend_module = sys.modules["fdtd.backend"]
_backend_module.backend_names = [dict(backends="numpy")]


def _keep_engine(name):
    return _backend_module.backend


fdtd_b200.set_backend = _backend_module.set_backend = _keep_engine
