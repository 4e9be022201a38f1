"""Frequency-domain post-processing of detector / source records (fdtd/fourier.py:15-260).

`FrequencyRoutines(grid, objs)` takes what the reference takes -- a plain array, a Line/Block
detector, a CurrentDetector, a SoftArbitraryPointSource, or a `(detector, current_detector)`
pair -- and offers the same three steps: `compute_padding`, `compute_frequencies`, `FFT`,
`impedance`.  The records come from the device ring buffers (detectors.py); the transforms run
as cuFFT calls on the engine's device in float64/complex128 (the reference's numpy backend
precision) and come back as host numpy arrays.

Kept quirks of the reference, so that scripts see the same numbers:
  * `FFT` of a detector transforms `E[:][0][0][0]`, i.e. the first recorded sample's [0][0]
    row, not a time trace (fdtd/fourier.py:186-195, marked FIXME there);
  * the default window drops the last bin (`end index -1`, fdtd/fourier.py:131-134);
  * `compute_padding` with a bin resolution but no bin count uses (begin - end) / resolution
    (fdtd/fourier.py:98-99).
One difference: a `freq_window_tuple` in `compute_frequencies` selects the closest bins as the
reference intends; the reference itself raises NameError there (fdtd/fourier.py:137-138).
"""
from math import ceil

import numpy as np
import torch

from .backend import backend as bd
from .detectors import BlockDetector, CurrentDetector, LineDetector
from .sources import SoftArbitraryPointSource


def _device_array(values):
    """nested lists / numpy / tensor -> float64 tensor on the engine's device."""
    if torch.is_tensor(values):
        return values.to(device=bd.device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(values, dtype=np.float64), device=bd.device)


def _pad_edge(t, n):
    """np.pad(t, (0, n), 'edge') for every axis, as numpy applies a single (before, after) pair."""
    if n <= 0:
        if n < 0:
            raise ValueError("index can't contain negative values")
        return t
    for axis in range(t.dim()):
        last = t.narrow(axis, t.shape[axis] - 1, 1)
        reps = [1] * t.dim()
        reps[axis] = n
        t = torch.cat([t, last.repeat(*reps)], dim=axis)
    return t


class FrequencyRoutines:
    verbose = True          # the reference prints the two resolutions on every call

    def __init__(self, grid, objs):
        self.grid = grid
        self.objs = objs

    # ------------------------------------------------------------------------------ helpers
    def compute_padding(self, input_data, dt, freq_window_tuple=None, fft_num_bins_in_window=None,
                        fft_bin_freq_resolution=None):
        """samples to append so that the window [begin, end] holds the requested number of FFT bins;
        returns (required_padding, end_time).  The padding is not applied here (fdtd/fourier.py:53-115)."""
        n = input_data.shape[0]
        if freq_window_tuple is None:
            lo, hi = 0, (n / 2.0) / (dt * n)
        else:
            lo, hi = freq_window_tuple
        end_time = n * dt
        bins = fft_num_bins_in_window
        if bins is None and fft_bin_freq_resolution is not None:
            bins = (lo - hi) / fft_bin_freq_resolution
        elif not (bins or fft_bin_freq_resolution):
            bins = n
        if self.verbose:
            print("Waveform data has an intrinsic resolution of: {:.2E} Hz".format(1.0 / end_time))
            print("FFT bin: {:.2E} Hz".format((1.0 / dt) / bins))
        return ceil(bins / ((hi - lo) * dt)) - n, end_time

    def compute_frequencies(self, length_with_padding, dt, freq_window_tuple=None):
        """(bin frequencies, first index, last index) of the window (fdtd/fourier.py:124-140)."""
        freqs = np.fft.fftfreq(length_with_padding, d=dt)
        if freq_window_tuple is None:
            return freqs, 0, -1
        lo, hi = freq_window_tuple
        return freqs, int(np.abs(freqs - lo).argmin()), int(np.abs(freqs - hi).argmin())

    def S_parameters(self, waveform=None, node=None):
        raise NotImplementedError

    def export_touchstone_s2p(self):
        raise NotImplementedError

    # ---------------------------------------------------------------------------- transforms
    def _record(self):
        o = self.objs
        if bd.is_array(o):
            return o
        if isinstance(o, CurrentDetector):
            return o.I[:][0][0][0]
        if isinstance(o, (LineDetector, BlockDetector)):
            return o.E[:][0][0][0]
        if isinstance(o, SoftArbitraryPointSource):
            return o.source_voltage[:][0][0][0]
        raise ValueError("Sorry, FFT can't yet interpret the argument given.")

    def FFT(self, freq_window_tuple=None, fft_num_bins_in_window=None, fft_bin_freq_resolution=None):
        """(frequencies, spectrum) of the record (fdtd/fourier.py:172-213)."""
        if self.grid.time_steps_passed == 0:
            return [], []
        data = _device_array(self._record())
        pad, _ = self.compute_padding(data, self.grid.time_step, freq_window_tuple=freq_window_tuple,
                                      fft_num_bins_in_window=fft_num_bins_in_window,
                                      fft_bin_freq_resolution=fft_bin_freq_resolution)
        data = _pad_edge(data, pad)
        spectrum = torch.fft.fft(data).cpu().numpy()
        freqs, a, b = self.compute_frequencies(data.shape[0], self.grid.time_step, freq_window_tuple=freq_window_tuple)
        return freqs[a:b], spectrum[a:b]

    def impedance(self, freq_window_tuple=None, fft_num_bins_in_window=None, fft_bin_freq_resolution=None):
        """(frequencies, V(f) / I(f)) of one node: a SoftArbitraryPointSource (its recorded output voltage and
        the current through its cell) or a (detector, current_detector) pair (fdtd/fourier.py:217-260)."""
        if self.grid.time_steps_passed == 0:
            return [], []
        o = self.objs
        if isinstance(o, SoftArbitraryPointSource):
            voltage = _device_array(o.source_voltage)
            current = _device_array(o.current_detector.I)
        elif (isinstance(o, tuple) and isinstance(o[0], (BlockDetector, LineDetector))
              and isinstance(o[1], CurrentDetector)):
            voltage = _device_array(o[0].E)[:][0, 0, 0]
            current = _device_array(o[1].I)
        else:
            raise ValueError("Sorry, FFT can't yet interpret the argument given.")
        pad, _ = self.compute_padding(voltage, self.grid.time_step, freq_window_tuple=freq_window_tuple,
                                      fft_num_bins_in_window=fft_num_bins_in_window,
                                      fft_bin_freq_resolution=fft_bin_freq_resolution)
        steps = self.grid.time_steps_passed
        voltage = _pad_edge(voltage.reshape(steps), pad)
        current = _pad_edge(current.reshape(steps), pad)
        v_f, i_f = torch.fft.fft(voltage), torch.fft.fft(current)
        safe = torch.where(i_f != 0, i_f, torch.ones_like(i_f))
        z_f = torch.where(i_f != 0, v_f / safe, torch.zeros_like(v_f)).cpu().numpy()
        freqs, a, b = self.compute_frequencies(voltage.shape[0], self.grid.time_step,
                                               freq_window_tuple=freq_window_tuple)
        return freqs[a:b], z_f[a:b]
