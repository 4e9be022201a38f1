"""Switch fdtd_b200's host layer onto the serial-interpreter build of its kernels (CPU tests only)."""
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
    backend._override_for_tests(_lib, "cpu", getattr(torch, dtype))
    return fdtd_b200
