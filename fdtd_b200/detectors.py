"""Detectors: LineDetector and BlockDetector (fdtd/detectors.py:20-280).

The reference appends one array per half-step to Python lists (`detector.E`, `detector.H`).
Here every half-step a small kernel gathers the detector's cells into a device ring buffer
`[capacity][points][3]`; the ring is copied to the host in batches (when it fills up, or when
the user reads the detector) and `detector.E` / `.H` materialise the reference's
list-of-per-step-arrays lazily from those host chunks.

Beyond the reference: `detector.track_frequencies(freqs)` keeps a running DFT of the record on
the device (one small kernel per ring flush, include/fdtd_b200.h `fdtd_dft_accumulate`), so the
spectrum of a long run on a large detector needs neither the host lists nor a device->host copy
of the time trace.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from .backend import backend as bd
from .sources import _diagonal_points, bounding_box, local_points


class _Detector:
    _kind = _capi.DET_FIELD
    _fields = ("E", "H")         # histories kept, in ring order (ring_E, ring_H)
    _width = 3                   # values per point and sample

    def __init__(self, name=None):
        self.grid = None
        self.name = name
        self._chunks = {"E": [], "H": []}     # host arrays (T_chunk, *sample_shape, 3)
        self._lists = {"E": None, "H": None}  # cached list view of the chunks
        self._ring_E = self._ring_H = None
        self._capacity = 0
        self._n_seen = {"E": 0, "H": 0}       # samples drained so far = record index of the next one
        self._dft_freqs = None
        self._dft_acc = {}
        self._keep_trace = True

    def _attach(self, grid):
        self.grid = grid
        self.grid.detectors.append(self)
        grid._register_name(self)

    def _set_points(self, xs, ys, zs, sample_shape):
        """global point list (sampling order) -> local subset on this rank."""
        self._sample_shape = tuple(sample_shape)
        self._n_points = int(np.prod(sample_shape))
        mine, lin = local_points(self.grid, xs, ys, zs)      # sorted by linear index
        self._n_local = len(lin)
        self._idx = torch.as_tensor(lin, dtype=torch.int64, device=bd.device)
        # ring column of each sorted entry = its rank among this slab's points in sampling order
        order = np.argsort(mine, kind="stable")
        pos = np.empty(len(mine), dtype=np.int32)
        pos[order] = np.arange(len(mine), dtype=np.int32)
        self._pos = torch.as_tensor(pos, dtype=torch.int32, device=bd.device)
        self._positions = mine[order]                         # global list positions, sampling order
        self._bbox = bounding_box(self.grid, lin)
        # which list positions every rank owns (the partition is known to all): lets a flush be ONE fixed-size
        # all_gather of the ring chunks instead of pickled objects
        part = self.grid._part
        if part.sharded:
            gx = np.asarray(xs, dtype=np.int64)
            gx = np.where(gx < 0, gx + self.grid.Nx, gx)
            self._rank_positions = [np.nonzero((gx >= part.bounds(r)[0]) & (gx < part.bounds(r)[1]))[0]
                                    for r in range(part.world)]
        # points of the rank that holds most of them: what every rank sizes the ring capacity with
        self._rank_pos_dev = None
        self._n_ring = max(len(p) for p in self._rank_positions) if part.sharded else self._n_local

    def _ensure_ring(self, capacity):
        if self._ring_E is None or self._capacity != capacity:
            g = self.grid
            if g._ring_fill["E"] or g._ring_fill["H"]:
                raise RuntimeError("detector ring resized with samples pending")
            self._capacity = capacity
            self._ring_E = bd.zeros((capacity, max(1, self._n_local), self._width), dtype=g._sdtype)
            self._ring_H = bd.zeros((capacity, max(1, self._n_local), self._width), dtype=g._sdtype)

    # ------------------------------------------------------------------ running DFT on the device
    def track_frequencies(self, frequencies, keep_trace=True):
        """From now on accumulate X(f) = sum_n x[n] exp(-2 pi i f n dt) of this detector's record on the device,
        n being the index of the sample in the record (so that, tracked from the first step at f = k / (N dt),
        it equals bin k of `numpy.fft.fft` over the N recorded samples).  `keep_trace=False` stops the time
        trace from being copied to the host at all: `detector.E` then stays empty."""
        if self.grid is None:
            raise RuntimeError("register the detector on a grid first")
        freqs = np.atleast_1d(np.asarray(frequencies, dtype=np.float64)).copy()
        if freqs.ndim != 1 or freqs.size == 0:
            raise ValueError("frequencies: a non-empty 1-D sequence in Hz")
        self._dft_freqs = freqs
        self._keep_trace = bool(keep_trace)
        self._dft_acc = {f: bd.zeros((freqs.size, max(1, self._n_local) * self._width, 2), dtype=torch.float64)
                         for f in self._chunks}

    @property
    def frequencies(self):
        return None if self._dft_freqs is None else self._dft_freqs.copy()

    def _accumulate(self, f, ring, n):
        """add ring rows [0, n) to the running DFT of field `f` (record indices _n_seen[f] ...)."""
        if self._n_local == 0:
            return
        import ctypes as C
        g = self.grid
        index = np.arange(self._n_seen[f], self._n_seen[f] + n, dtype=np.float64)
        phase = (-2.0 * np.pi) * ((index * g.time_step)[:, None] * self._dft_freqs[None, :])
        tw = torch.as_tensor(np.stack([np.cos(phase), np.sin(phase)], axis=-1), device=ring.device)
        lib = bd.lib
        rc = lib.fdtd_dft_accumulate(_capi.F32 if ring.dtype is torch.float32 else _capi.F64,   # (storage type)
                                     C.c_void_p(ring.data_ptr()), n, self._n_local * self._width,
                                     C.c_void_p(tw.data_ptr()), self._dft_freqs.size,
                                     C.c_void_p(self._dft_acc[f].data_ptr()), g._engine._stream())
        _capi.check(lib, rc)

    def spectrum(self, field="E"):
        """complex array (n_frequencies, *sample_shape[, 3]) of the tracked DFT of `field`."""
        if self._dft_freqs is None:
            raise RuntimeError("call track_frequencies(...) before running")
        if field not in self._dft_acc:
            raise KeyError(field)
        g = self.grid
        if g._engine is not None:
            g._engine.flush_detectors()
        nf = self._dft_freqs.size
        acc = self._dft_acc[field].view(nf, max(1, self._n_local), self._width, 2)
        part = g._part
        if part.sharded:
            n_max = max(1, max(len(p) for p in self._rank_positions))
            send = acc.new_zeros((nf, n_max, self._width, 2))
            send[:, :self._n_local] = acc[:, :self._n_local]
            recv = [torch.empty_like(send) for _ in range(part.world)]
            dist.all_gather(recv, send)
            recv = torch.stack(recv).to("cpu").numpy()
            full = np.zeros((nf, self._n_points, self._width, 2))
            for r, pos in enumerate(self._rank_positions):
                full[:, pos] = recv[r, :, :len(pos)]
        else:
            full = acc[:, :self._n_local].to("cpu", copy=True).numpy()
        out = full[..., 0] + 1j * full[..., 1]
        tail = (self._width,) if self._width > 1 else ()
        return out.reshape((nf,) + self._sample_shape + tail)

    @property
    def spectrum_E(self):
        return self.spectrum("E")

    @property
    def spectrum_H(self):
        return self.spectrum("H")

    def _drain(self, nE, nH):
        """copy the filled part of the rings to the host (one batch) and assemble global samples.

        x-sharded grids: ONE all_gather carries the rows of both fields (padded to the rank that holds most points),
        the global samples are assembled on the device and only they cross to the host -- not the padding of every
        rank."""
        part = self.grid._part
        todo = [(f, ring, n) for f, ring, n in (("E", self._ring_E, nE), ("H", self._ring_H, nH))
                if n and f in self._chunks]
        for f, ring, n in todo:
            if self._dft_freqs is not None:
                self._accumulate(f, ring, n)
            self._n_seen[f] += n
        if not self._keep_trace or not todo:
            return
        tail = (self._width,) if self._width > 1 else ()
        if part.sharded:
            n_max = max(1, max(len(p) for p in self._rank_positions))
            rows = sum(n for _, _, n in todo)
            ring0 = todo[0][1]
            send = ring0.new_zeros((rows, n_max, self._width))
            at = 0
            for _, ring, n in todo:
                send[at:at + n, :self._n_local] = ring[:n, :self._n_local]
                at += n
            recv = [torch.empty_like(send) for _ in range(part.world)]
            dist.all_gather(recv, send)
            if getattr(self, "_rank_pos_dev", None) is None:
                self._rank_pos_dev = [torch.as_tensor(np.asarray(pos, dtype=np.int64), device=send.device)
                                      for pos in self._rank_positions]
            full = send.new_zeros((rows, self._n_points, self._width))
            for r, pos in enumerate(self._rank_pos_dev):
                if pos.numel():
                    full[:, pos] = recv[r][:, :pos.numel()]
            full = full.to("cpu").numpy()
            at = 0
            for f, _, n in todo:
                self._chunks[f].append(full[at:at + n].reshape((n,) + self._sample_shape + tail))
                self._lists[f] = None
                at += n
            return
        for f, ring, n in todo:
            full = ring[:n, :self._n_local].to("cpu", copy=True).numpy()
            self._chunks[f].append(full.reshape((n,) + self._sample_shape + tail))
            self._lists[f] = None

    def _history(self, f):
        g = self.grid
        if g is not None and g._engine is not None:
            g._engine.flush_detectors()
        if self._lists[f] is None:
            self._lists[f] = [step for chunk in self._chunks[f] for step in chunk]
        return self._lists[f]

    @property
    def E(self):
        """list with one array per E half-step, as the reference's `detector.E`."""
        return self._history("E")

    @property
    def H(self):
        return self._history("H")

    def detector_values(self):
        """outputs what detector detects (fdtd/detectors.py:137-139)."""
        return {"E": self.E, "H": self.H}

    def detect_E(self):      # plug-in protocol of the reference; sampling happens on the device
        pass

    def detect_H(self):
        pass

    detect_E._fdtd_b200_builtin = detect_H._fdtd_b200_builtin = True

    def __repr__(self):
        return f"{self.__class__.__name__}(name={repr(self.name)})"

    def __str__(self):
        s = "    " + repr(self) + "\n"
        return s + (f"        @ x=[{self.x[0]}, ... , {self.x[-1]}], y=[{self.y[0]}, ... , {self.y[-1]}], "
                    f"z=[{self.z[0]}, ... , {self.z[-1]}]\n")


class LineDetector(_Detector):
    """samples E and H along the diagonal of a box: one (L, 3) array per half-step
    (fdtd/detectors.py:20-139)."""

    def _register_grid(self, grid, x, y, z):
        self._attach(grid)
        self.x, self.y, self.z = _diagonal_points(grid, x, y, z, False, 0, "LineDetector")
        self._set_points(self.x, self.y, self.z, (len(self.x),))


class BlockDetector(_Detector):
    """samples a block with INCLUSIVE upper bounds (fdtd/detectors.py:236-238): one
    (nx, ny, nz, 3) array per half-step, indexable like the reference's nested lists."""

    def _register_grid(self, grid, x, y, z):
        self._attach(grid)
        if isinstance(x, list) and isinstance(y, list) and isinstance(z, list):
            if len(x) != len(y) or len(y) != len(z) or len(z) != len(x):
                raise IndexError("sources require grid to be indexed with slices or equal length list-indices")
            self.x, self.y, self.z = x, y, z
        else:
            out = []
            for s, n in ((x, grid.Nx), (y, grid.Ny), (z, grid.Nz)):
                if isinstance(s, list):
                    s = slice(s[0], s[-1], None)
                a = s.start if s.start is not None else 0
                b = s.stop if s.stop is not None else n
                out.append(list(range(a, b + 1)))
            self.x, self.y, self.z = out
        for v, n in ((self.x, grid.Nx), (self.y, grid.Ny), (self.z, grid.Nz)):
            if any(i >= n or i < -n for i in v):
                # the reference fails with IndexError at the first step (SURVEY.md 8a trap 9)
                raise IndexError("BlockDetector ranges are inclusive of `stop`: index out of range")
        X, Y, Z = np.meshgrid(self.x, self.y, self.z, indexing="ij")
        self._set_points(X.ravel(), Y.ravel(), Z.ravel(), X.shape)


class CurrentDetector(BlockDetector):
    """z-directed current through each cell of a block (inclusive ranges like BlockDetector), from the
    loop of H around the cell averaged over two z levels (fdtd/detectors.py:284-496).  One (nx, ny, nz)
    array per step in `detector.I`, sampled on the device after every H half-step; the most recent sample
    also stays on the device, where a SoftArbitraryPointSource reads it back without a host round trip."""

    _kind = _capi.DET_CURRENT
    _width = 1

    def __init__(self, name=None):
        super().__init__(name)
        self.orientation = None
        self._chunks = {"H": []}          # the current is sampled with the H half-step
        self._lists = {"H": None}

    def _register_grid(self, grid, x, y, z):
        super()._register_grid(grid, x, y, z)
        part = grid._part
        self._needs_ghost = self._needs_wrap = False
        if part.sharded:
            xs = {(v + grid.Nx) % grid.Nx for v in self.x}
            # a cell on the first plane of a slab reads the left neighbour's H of the same half-step: the engine then
            # samples after the ghost plane has arrived (same verdict on every rank).  On global plane x = 0 that
            # neighbour is the LAST slab: H[x-1] = H[-1] wraps like python indexing (fdtd/detectors.py:432-447), and
            # the last slab's last plane travels into the first slab's low ghost plane before sampling.
            self._needs_wrap = 0 in xs
            self._needs_ghost = self._needs_wrap or bool(xs & {part.bounds(r)[0] for r in range(1, part.world)})
        self._last = bd.zeros((max(1, self._n_local),), dtype=grid._sdtype)

    @property
    def I(self):
        return self._history("H")

    @property
    def spectrum_I(self):
        return self.spectrum("H")

    @property
    def E(self):
        raise AttributeError("CurrentDetector records I, not E")

    @property
    def H(self):
        raise AttributeError("CurrentDetector records I, not H")

    spectrum_E = E
    spectrum_H = H

    def detector_values(self):
        """outputs what detector detects (fdtd/detectors.py:494-496)."""
        return {"I": self.I}
