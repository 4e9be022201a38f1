"""Optional conversions between simulation units and SI field units (fdtd/conversions.py:29-45).

In the engine's reduced units the free-space impedance is 1: E carries a factor sqrt(eps0) and H a factor
sqrt(mu0) relative to SI.  Works on numbers, numpy arrays and tensors alike."""
from math import sqrt

from . import constants as const

_SQRT_EPS0 = sqrt(const.eps0)
_SQRT_MU0 = sqrt(const.mu0)


def simE_to_worldE(input):
    return input / _SQRT_EPS0


def worldE_to_simE(input):
    return _SQRT_EPS0 * input


def simH_to_worldH(input):
    return input / _SQRT_MU0


def worldH_to_simH(input):
    return _SQRT_MU0 * input
