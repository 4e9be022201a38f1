"""Point fdtd_b200's host layer at the serial-interpreter build of its kernels (CPU tests only).

The product has no hook for this: the test pokes the backend singleton's attributes directly, the same
ones `fdtd_b200.set_backend` fills in on a GPU box (library handle, device, dtype)."""
import torch

import fdtd_b200
from fdtd_b200 import _capi
from fdtd_b200.backend import backend

from . import build_emu

_lib = None


def use_emu(dtype="float64"):
    global _lib
    if _lib is None:
        _lib = _capi.bind(build_emu.build())
    backend.lib = _lib
    backend.device = torch.device("cpu")
    # "float32x": float32 state, float64 arithmetic and coefficients (fdtd_b200/backend.py)
    backend.float = torch.float64 if dtype == "float32x" else getattr(torch, dtype)
    backend.storage = torch.float32 if dtype == "float32x" else backend.float
    backend.name = f"emu.{dtype}"
    return fdtd_b200
