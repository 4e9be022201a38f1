"""Waveforms evaluated on the host, in Python floats, exactly as the reference does
(fdtd/waveforms.py:8-9 and fdtd/sources.py:93-108, 278-295, 476-486): the per-step scalars
are tabulated on the host and uploaded, never recomputed on the device (SURVEY.md 8a trap 6)."""
from math import cos, pi, sin


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
