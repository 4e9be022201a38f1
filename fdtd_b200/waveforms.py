"""Waveforms evaluated on the host, in Python floats, exactly as the reference does
(fdtd/waveforms.py:8-9 and fdtd/sources.py:93-108, 278-295, 476-486): the per-step scalars
are tabulated on the host and uploaded, never recomputed on the device (SURVEY.md 8a trap 6)."""
from math import cos, exp, log, pi, sin, sqrt

import numpy as np


def hanning(f, t, n):
    """Hanning-windowed sine (fdtd/waveforms.py:8-9)."""
    return (1 / 2) * (1 - cos(f * t / n)) * (sin(f * t))


def continuous(q, period, phase_shift):
    """sin(2 pi q / period + phase) (fdtd/sources.py:108, 295, 479)."""
    return sin(2 * pi * q / period + phase_shift)


def pulse(q, frequency, hanning_dt, cycle):
    """the `pulse=True` branch of Point/LineSource (fdtd/sources.py:97-105, 282-292)."""
    t1 = int(2 * pi / (frequency * hanning_dt / cycle))
    if q < t1:
        return hanning(frequency, q * hanning_dt, cycle)
    return 0


# ---- pulse shapes for SoftArbitraryPointSource waveform arrays (fdtd/waveforms.py:33-52), peak value 1 --------
FWHM_PER_SIGMA = 2.0 * sqrt(2.0 * log(2))
fwhm_constant = FWHM_PER_SIGMA          # the reference's name


def normalized_gaussian_pulse(x, fwhm, center=0.0):
    """exp(-(x - center)^2 / (2 sigma^2)) with sigma = fwhm / (2 sqrt(2 ln 2)); x may be an array."""
    sigma = fwhm / FWHM_PER_SIGMA
    return np.exp(-(((x - center) ** 2.0) / (2.0 * (sigma ** 2.0))))


def normalized_gaussian_derivative_pulse(x, fwhm, center=0.0):
    """first derivative of the gaussian, scaled to a peak of 1 (scalar x, like the reference)."""
    sigma = fwhm / FWHM_PER_SIGMA
    return (exp((1.0 / 2.0) - ((x - center) ** 2.0) / (2.0 * sigma ** 2.0)) * (x - center)) / sigma
