"""Sources: PointSource, LineSource (soft, on Ez) and PlaneSource (hard, on one E and one H
component) with the reference's registration semantics (fdtd/sources.py:25-501).

Per step the reference evaluates one Python scalar per source (`math.sin` / `hanning`) and
writes it into the fields (fdtd/sources.py:93-109, 278-297, 476-486).  Here the scalars are
tabulated on the host with the same expressions, uploaded once per run, and injected on the
device; nothing per-step happens in Python.
"""
from math import pi, sin

import numpy as np
import torch

from . import _capi
from .backend import backend as bd
from ._hostmath import HostLib, scalar_in_dtype
from .waveforms import continuous, pulse


def _diagonal_points(grid, x, y, z, convert, min_points, what):
    """index lists of the diagonal through the box given by slices / lists
    (fdtd/sources.py:209-276, fdtd/detectors.py:62-112).  numpy linspace semantics."""
    conv = grid._handle_distance if convert else (lambda v: v)
    if isinstance(x, list) and isinstance(y, list) and isinstance(z, list):
        if len(x) != len(y) or len(y) != len(z) or len(z) != len(x):
            raise IndexError("sources require grid to be indexed with slices or equal length list-indices")
        return [conv(v) for v in x], [conv(v) for v in y], [conv(v) for v in z]
    ends = []
    for s, n in ((x, grid.Nx), (y, grid.Ny), (z, grid.Nz)):
        if isinstance(s, list):
            s = slice(conv(s[0]), conv(s[-1]), None)
        a = conv(s.start if s.start is not None else 0)
        b = conv(s.stop if s.stop is not None else n)
        ends.append((a, b))
    m = max(abs(b - a) for a, b in ends)
    if m < min_points:
        raise ValueError(f"a {what} should consist of at least two gridpoints")
    out = [[int(v) for v in np.linspace(a, b, m, endpoint=False).astype(np.int64)] for a, b in ends]
    return out[0], out[1], out[2]


def bounding_box(grid, lin):
    """local half-open bounding box x0,x1,y0,y1,z0,z1 of local linear cell indices."""
    if len(lin) == 0:
        return [0, 0, 0, 0, 0, 0]
    lin = np.asarray(lin, dtype=np.int64)
    plane = grid.Ny * grid.Nz
    x, y, z = lin // plane, (lin % plane) // grid.Nz, lin % grid.Nz
    return [int(x.min()), int(x.max()) + 1, int(y.min()), int(y.max()) + 1, int(z.min()), int(z.max()) + 1]


def local_points(grid, xs, ys, zs):
    """-> (positions in the list, local linear cell indices) of the points this rank owns,
    sorted by ascending linear index (the C ABI wants ascending `idx`)."""
    part = grid._part
    xs, ys, zs = (np.asarray(v, dtype=np.int64) for v in (xs, ys, zs))
    # negative indices address from the end, as numpy / torch indexing does in the reference
    xs = np.where(xs < 0, xs + grid.Nx, xs)
    ys = np.where(ys < 0, ys + grid.Ny, ys)
    zs = np.where(zs < 0, zs + grid.Nz, zs)
    if ((xs < 0) | (xs >= grid.Nx) | (ys < 0) | (ys >= grid.Ny) | (zs < 0) | (zs >= grid.Nz)).any():
        raise IndexError("index out of range for the grid")
    mine = np.nonzero((xs >= part.x0) & (xs < part.x1))[0]
    lin = (xs[mine] - part.x0) * (grid.Ny * grid.Nz) + ys[mine] * grid.Nz + zs[mine]
    order = np.argsort(lin, kind="stable")
    return mine[order], lin[order]


class _TimedSource:
    """period / pulse bookkeeping shared by Point and Line sources."""

    def __init__(self, period, amplitude, phase_shift, name, pulse, cycle, hanning_dt):
        self.grid = None
        self.period = period
        self.amplitude = amplitude
        self.phase_shift = phase_shift
        self.name = name
        self.pulse = pulse
        self.cycle = cycle
        self.frequency = 1.0 / period
        self.hanning_dt = hanning_dt if hanning_dt is not None else 0.5 / self.frequency

    def _attach(self, grid):
        self.grid = grid
        self.grid.sources.append(self)
        grid._register_name(self)

    def _scalar(self, q):
        if self.pulse:
            return pulse(q, self.frequency, self.hanning_dt, self.cycle)
        return continuous(q, self.period, self.phase_shift)

    def _signature(self):
        return (self.period, self.amplitude, self.phase_shift, self.pulse, self.cycle, self.frequency, self.hanning_dt)

    def update_E(self):      # the reference's plug-in protocol; the engine does the work
        pass

    def update_H(self):
        pass

    update_E._fdtd_b200_builtin = update_H._fdtd_b200_builtin = True

    def __repr__(self):
        return (f"{self.__class__.__name__}(period={self.period}, amplitude={self.amplitude}, "
                f"phase_shift={self.phase_shift}, name={repr(self.name)})")


class PointSource(_TimedSource):
    """A source placed at a single grid cell: Ez += amplitude * waveform(q) (fdtd/sources.py:25-127)."""

    def __init__(self, period=15, amplitude: float = 1.0, phase_shift: float = 0.0, name: str = None,
                 pulse: bool = False, cycle: int = 5, hanning_dt: float = 10.0):
        super().__init__(period, amplitude, phase_shift, name, pulse, cycle, hanning_dt)

    def _register_grid(self, grid, x, y, z):
        self._attach(grid)
        try:
            (x,), (y,), (z,) = x, y, z
        except (TypeError, ValueError):
            raise ValueError("a point source should be placed on a single grid cell.")
        self.x, self.y, self.z = grid._handle_tuple((x, y, z))
        self.period = grid._handle_time(self.period)
        self.frequency = 1.0 / self.period
        _, lin = local_points(grid, [self.x], [self.y], [self.z])
        self._idx = torch.as_tensor(lin, dtype=torch.int64, device=bd.device)
        self._bbox = bounding_box(grid, lin)
        self._profile = bd.ones((len(lin),))

    def _wave_value(self, q):
        # the reference adds the Python double amplitude*waveform to the array element
        return self.amplitude * self._scalar(q)

    def _entries(self):
        return [dict(kind=_capi.SRC_POINTS, field=0, comp=2, idx=self._idx, profile=self._profile, bbox=self._bbox)]

    def __str__(self):
        return "    " + repr(self) + "\n" + f"        @ x={self.x}, y={self.y}, z={self.z}\n"


class LineSource(_TimedSource):
    """A source along the diagonal of a box, gaussian profile, on Ez (fdtd/sources.py:131-315)."""

    def __init__(self, period=15, amplitude: float = 1.0, phase_shift: float = 0.0, name: str = None,
                 pulse: bool = False, cycle: int = 5, hanning_dt: float = 10.0):
        super().__init__(period, amplitude, phase_shift, name, pulse, cycle, hanning_dt)

    def _register_grid(self, grid, x, y, z):
        self._attach(grid)
        self.x, self.y, self.z = _diagonal_points(grid, x, y, z, True, 2, "LineSource")
        self.period = grid._handle_time(self.period)
        self.frequency = 1.0 / self.period
        # gaussian profile over the SQUARED distance to the middle point (fdtd/sources.py:197-207)
        hl = HostLib(grid._dtype)
        L = len(self.x)
        ix, iy, iz = (np.array(v) for v in (self.x, self.y, self.z))
        vect = hl.asarray((ix - self.x[L // 2]) ** 2 + (iy - self.y[L // 2]) ** 2 + (iz - self.z[L // 2]) ** 2)
        profile = hl.exp(-(vect ** 2) / (2 * (0.5 * vect.max()) ** 2))
        profile /= profile.sum()
        profile *= self.amplitude
        self.profile = hl.to_device(profile, bd.device)
        mine, lin = local_points(grid, self.x, self.y, self.z)
        self._idx = torch.as_tensor(lin, dtype=torch.int64, device=bd.device)
        self._bbox = bounding_box(grid, lin)
        self._profile = self.profile[torch.as_tensor(mine, dtype=torch.int64, device=bd.device)].contiguous()

    def _wave_value(self, q):
        return self._scalar(q)

    def _entries(self):
        return [dict(kind=_capi.SRC_POINTS, field=0, comp=2, idx=self._idx, profile=self._profile, bbox=self._bbox)]

    def __str__(self):
        s = "    " + repr(self) + "\n"
        return s + (f"        @ x=[{self.x[0]}, ... , {self.x[-1]}], y=[{self.y[0]}, ... , {self.y[-1]}], "
                    f"z=[{self.z[0]}, ... , {self.z[-1]}]\n")


class PlaneSource:
    """A hard source on a one-cell-thick plane: E[pol] = H[hpol] = amplitude*sin(2 pi q/period + phase)
    (fdtd/sources.py:319-501)."""

    def __init__(self, period=15, amplitude: float = 1.0, phase_shift: float = 0.0, name: str = None,
                 polarization: str = "z"):
        self.grid = None
        self.period = period
        self.amplitude = amplitude
        self.phase_shift = phase_shift
        self.name = name
        self.polarization = polarization

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        self.x, self.y, self.z = self._handle_slices(x, y, z)
        self._resolved()                   # out-of-range regions fail here, not at the first step, and register nothing
        self.grid.sources.append(self)
        grid._register_name(self)
        self.period = grid._handle_time(self.period)
        self.frequency = 1.0 / self.period
        ext = [s.stop - s.start for s in (self.x, self.y, self.z)]
        self.profile = self.amplitude * bd.ones(tuple(ext))      # reference attribute (uniform)
        self._amplitude0 = self.amplitude     # the reference bakes the amplitude into `profile` here, once

    def _handle_slices(self, x, y, z):
        g = self.grid
        out = []
        for s, n in ((x, g.Nx), (y, g.Ny), (z, g.Nz)):
            if not isinstance(s, slice):
                if isinstance(s, list):
                    (s,) = s
                s = slice(g._handle_distance(s), g._handle_distance(s) + 1, None)
            a = g._handle_distance(s.start if s.start is not None else 0)
            b = g._handle_distance(s.stop if s.stop is not None else n)
            out.append(slice(a, b) if a < b else (slice(b, a) if a > b else slice(a, a + 1)))
        x, y, z = out
        ext = [x.stop - x.start, y.stop - y.start, z.stop - z.start]
        if ext.count(0) > 0:
            raise ValueError("Given location for PlaneSource results in slices of length 0!")
        if ext.count(1) == 0:
            raise ValueError("Given location for PlaneSource is not a 2D plane!")
        if ext.count(1) > 1:
            raise ValueError("Given location for PlaneSource should have no more than one dimension "
                             "in which it's flat.\nUse a LineSource for lower dimensional sources.")
        self._Epol = "xyz".index(self.polarization)
        if ext[self._Epol] == 1:
            raise ValueError("PlaneSource cannot be polarized perpendicular to the orientation of the plane.")
        # H component: fdtd/sources.py:468-472 (pinned by the reference's tests/test_sources.py:25-43)
        probe, first, second = [(2, 1, 2), (2, 0, 2), (1, 0, 1)][self._Epol]
        self._Hpol = first if ext[probe] == 1 else second
        return x, y, z

    def _wave_value(self, q):
        return sin(2 * pi * q / self.period + self.phase_shift)

    def _signature(self):
        return (self.period, self.phase_shift)

    def _resolved(self):
        """the cells `E[self.x, self.y, self.z]` addresses in the reference: slice bounds resolved the way numpy
        resolves them (negative bounds count from the end).  A region the `profile` (shaped by the nominal extents)
        could not be assigned to raises, as the reference's assignment would at the first step."""
        g = self.grid
        out = []
        for s, n in ((self.x, g.Nx), (self.y, g.Ny), (self.z, g.Nz)):
            a, b, _ = slice(s.start, s.stop).indices(n)
            b = max(a, b)
            nominal = s.stop - s.start
            if b - a != nominal and not (nominal == 1 and b == a):     # (1 -> 0 broadcasts: a silent no-op there)
                raise IndexError(f"PlaneSource region {s.start}:{s.stop} reaches outside the grid (extent {n})")
            out.append((a, b))
        return out

    def _entries(self):
        g = self.grid
        (x0, x1), (y0, y1), (z0, z1) = self._resolved()
        lx0, lx1 = g._part.local_range(x0, x1)
        box = [lx0, lx1, y0, y1, z0, z1]
        amp = scalar_in_dtype(self._amplitude0, g._dtype)
        return [dict(kind=_capi.SRC_BOX, field=0, comp=self._Epol, box=box, amplitude=amp),
                dict(kind=_capi.SRC_BOX, field=1, comp=self._Hpol, box=box, amplitude=amp)]

    def update_E(self):
        pass

    def update_H(self):
        pass

    update_E._fdtd_b200_builtin = update_H._fdtd_b200_builtin = True

    def __repr__(self):
        return (f"{self.__class__.__name__}(period={self.period}, amplitude={self.amplitude}, "
                f"phase_shift={self.phase_shift}, name={repr(self.name)}, "
                f"polarization={repr(self.polarization)})")

    def __str__(self):
        s = "    " + repr(self) + "\n"
        return s + (f"        @ x=[{self.x.start}, ... , {self.x.stop}], y=[{self.y.start}, ... , {self.y.stop}], "
                    f"z=[{self.z.start}, ... , {self.z.stop}]\n")


class SoftArbitraryPointSource:
    """A voltage source with series impedance on one cell, paired with a CurrentDetector on the same cell
    (fdtd/sources.py:504-643):  Ez += (waveform[q] + Z * I_prev) / grid_spacing, where I_prev is the
    detector's sample of the previous step.  The reference reads that sample back on the host every step;
    here it stays on the device (the detector kernel leaves its latest value in a device scalar the source
    kernel reads), so a 50-ohm feed costs no host round trip.  `input_voltage` / `source_voltage` are
    recorded in a device ring and materialised like detector histories."""

    def __init__(self, waveform_array, name: str = None, impedance: float = 0.0):
        self.grid = None
        self.name = name
        self.current_detector = None
        self.waveform_array = waveform_array
        self.impedance = impedance
        self._chunks = []
        self._lists = None
        self._ring_V = None
        self._capacity = 0
        self._steps_logged = []      # step index of every E half-step this source took part in
        self._wf_cache = None

    def _register_grid(self, grid, x, y, z):
        from .detectors import CurrentDetector
        self.grid = grid
        self.grid.sources.append(self)
        grid._register_name(self)
        try:
            (x,), (y,), (z,) = x, y, z
        except (TypeError, ValueError):
            raise ValueError("a point source should be placed on a single grid cell.")
        self.x, self.y, self.z = grid._handle_tuple((x, y, z))
        # (the reference raises UnboundLocalError here when a name is given, fdtd/sources.py:592-593)
        detector_name = self.name + "_I" if self.name is not None else None
        self.current_detector = CurrentDetector(name=detector_name)
        grid[x, y, z] = self.current_detector
        # negative indices address from the end, as the reference's `grid.E[x, y, z, 2] += ...` does
        cell = []
        for v, n in ((self.x, grid.Nx), (self.y, grid.Ny), (self.z, grid.Nz)):
            c = v + n if v < 0 else v
            if not 0 <= c < n:
                raise IndexError(f"index {v} is out of bounds for a grid axis of size {n}")
            cell.append(c)
        lx0, lx1 = grid._part.local_range(cell[0], cell[0] + 1)
        self._n_local = lx1 - lx0
        self._box = [lx0, lx1, cell[1], cell[1] + 1, cell[2], cell[2] + 1]

    @property
    def _waveform_host(self):
        """the waveform as a flat host array (re-read when the user assigns a new `waveform_array`)"""
        wf = self.waveform_array
        if self._wf_cache is None or self._wf_cache[0] is not wf:
            arr = wf.detach().cpu().numpy() if torch.is_tensor(wf) else np.asarray(wf)
            self._wf_cache = (wf, arr.reshape(-1))
        return self._wf_cache[1]

    def _wave_value(self, q):
        # input voltage of step q; zero once the waveform is exhausted (fdtd/sources.py:601-605)
        return float(self._waveform_host[q]) if q < self._waveform_host.shape[0] else 0.0

    def _entries(self):
        return [dict(kind=_capi.SRC_FEEDBACK, field=0, comp=2, n=self._n_local, box=self._box,
                     impedance=float(self.impedance), feedback=self.current_detector._last)]

    def _signature(self):
        return (id(self.waveform_array), float(self.impedance))

    def _ensure_ring(self, capacity):
        if self._ring_V is None or self._capacity != capacity:
            self._capacity = capacity
            self._ring_V = bd.zeros((capacity,), dtype=self.grid._sdtype)

    def _drain(self, n):
        if n == 0:
            return
        local = self._ring_V[:n].to("cpu", copy=True).numpy()
        part = self.grid._part
        if part.sharded:
            import torch.distributed as dist
            gathered = [None] * part.world
            dist.all_gather_object(gathered, local if self._n_local else None)
            local = next(v for v in gathered if v is not None)
        self._chunks.append(local)
        self._lists = None

    def _history(self):
        g = self.grid
        if g is not None and g._engine is not None:
            g._engine.flush_detectors()
        if self._lists is None:
            vout = [v for chunk in self._chunks for v in chunk]
            # the input voltage is host data (the reference records the waveform element itself)
            vin = [self._waveform_host[q] if q < self._waveform_host.shape[0] else 0.0
                   for q in self._steps_logged[:len(vout)]]
            if not self.impedance > 0:
                vout = vin                       # no feedback: the output voltage is the waveform element
            self._lists = ([[[[v]]] for v in vin], [[[[v]]] for v in vout])
        return self._lists

    @property
    def input_voltage(self):
        """voltage hard-imposed by the source, one [[[v]]] per step like the reference's list"""
        return self._history()[0]

    @property
    def source_voltage(self):
        return self._history()[1]

    def update_E(self):
        pass

    def update_H(self):
        pass

    update_E._fdtd_b200_builtin = update_H._fdtd_b200_builtin = True

    def __repr__(self):
        return f"{self.__class__.__name__}()"

    def __str__(self):
        return "    " + repr(self) + "\n" + f"        @ x={self.x}, y={self.y}, z={self.z}\n"
