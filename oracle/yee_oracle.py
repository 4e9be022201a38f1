"""CPU oracle for the per-timestep Yee update of flaport/fdtd.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement of the reference's
hot path, used as the checker in ``tests/``, in ``__graft_entry__.smoke()`` and
as the timed ``cpu_baseline`` / ``--impl reference`` leg of ``bench.py``.  The
product (``fdtd_b200``) never imports it.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this file against
  * the reference's own golden curl vectors (reference tests/test_grid.py:47-164),
  * ``tests/golden/*.npz`` -- outputs of the unmodified reference (numpy float64
    and true-float32 torch) produced in the build container by
    ``tests/golden/make_golden.py``; E, H and every detector trace are
    bit-identical for every committed scene.

What is restated (reference file:line):
  curl_E / curl_H ............ fdtd/grid.py:29-76
  Grid.__init__ / indices .... fdtd/grid.py:90-223, 358-375
  Grid.update_E / update_H ... fdtd/grid.py:275-325   (ordering of the sub-steps)
  PML coefficients ........... fdtd/boundaries.py:291-293, 367-407, 490-625
  PML psi / phi / add ........ fdtd/boundaries.py:409-487
  PeriodicBoundary ........... fdtd/boundaries.py:143-219
  Object / Absorbing / Aniso . fdtd/objects.py:37-129, 163-221, 232-269
  Point/Line/PlaneSource ..... fdtd/sources.py:25-127, 131-315, 319-501
  hanning .................... fdtd/waveforms.py:8-9
  Line/BlockDetector ......... fdtd/detectors.py:20-139, 146-280
  CurrentDetector ............ fdtd/detectors.py:284-496
  SoftArbitraryPointSource ... fdtd/sources.py:504-643

It is NOT a copy: the reference keeps nine psi components, three phi components
and six full-size coefficient arrays per PML and loops over plug-in objects; the
restatement keeps the two psi scalars per slab cell that can ever be non-zero
and 1-D b/c profiles ("verified fused restatement").  The
arithmetic per cell -- every product, sum and their order -- is the reference's,
which is why the results are bit-identical.

The array library is pluggable (numpy, or torch-CPU for the multi-threaded
baseline): only slicing and elementwise + - * / are used.
"""
from math import pi, sin, cos
import numpy as _np

C0 = 299792458.0            # fdtd/constants.py:5
MU0 = 4e-7 * pi             # fdtd/constants.py:17
ETA0 = MU0 * C0             # fdtd/constants.py:23


# --------------------------------------------------------------------------- array shim
class _Lib:
    """Minimal array interface shared by the numpy and torch flavours."""

    def __init__(self, kind="numpy", dtype="float64", device=None):
        self.kind = kind
        self.kw = {}
        if kind == "numpy":
            self.mod = _np
            self.dtype = getattr(_np, dtype)
        elif kind == "torch":
            import torch
            self.mod = torch
            self.dtype = getattr(torch, dtype)
            # device="cuda": the same slicing code as eager ATen kernels on a GPU, which is how the reference's
            # `torch.cuda` backends run (fdtd/backend.py:322-355) -- bench.py's `gpu_eager_baseline`, nothing else
            if device is not None:
                self.kw = {"device": device}
        else:
            raise ValueError(kind)

    def zeros(self, shape):
        return self.mod.zeros(tuple(shape), dtype=self.dtype, **self.kw)

    def ones(self, shape):
        return self.mod.ones(tuple(shape), dtype=self.dtype, **self.kw)

    def asarray(self, a):
        if self.kind == "numpy":
            return _np.array(a, dtype=self.dtype)
        import torch
        if torch.is_tensor(a):
            return a.clone().to(dtype=self.dtype, **self.kw)
        return torch.tensor(_np.asarray(a), dtype=self.dtype, **self.kw)

    def is_array(self, a):
        if isinstance(a, _np.ndarray):
            return True
        if self.kind == "torch":
            import torch
            return torch.is_tensor(a)
        return False

    def arange(self, a, b, s):
        if self.kind == "numpy":
            return _np.asarray(_np.arange(a, b, s), dtype=self.dtype)
        return self.mod.arange(a, b, s, dtype=self.dtype, **self.kw)

    def exp(self, a):
        return self.mod.exp(a)

    def copy(self, a):
        return a.copy() if self.kind == "numpy" else a.clone()

    def to_numpy(self, a):
        return a if isinstance(a, _np.ndarray) else a.cpu().numpy()


lib = _Lib()


def set_backend(kind="numpy", dtype="float64", device=None):
    """Choose the oracle's array library and precision (before building a grid)."""
    global lib
    lib = _Lib(kind, dtype, device)
    return lib


# --------------------------------------------------------------------------- curls
def _bdiff(F, axis):
    """backward difference along `axis`, zero at index 0 (fdtd/grid.py:66-74)."""
    out = lib.zeros(F.shape)
    n = F.shape[axis]
    if n > 1:
        hi = [slice(None)] * 3
        lo = [slice(None)] * 3
        hi[axis] = slice(1, None)
        lo[axis] = slice(None, -1)
        out[tuple(hi)] = F[tuple(hi)] - F[tuple(lo)]
    return out


def _fdiff(F, axis):
    """forward difference along `axis`, zero at index N-1 (fdtd/grid.py:41-49)."""
    out = lib.zeros(F.shape)
    n = F.shape[axis]
    if n > 1:
        hi = [slice(None)] * 3
        lo = [slice(None)] * 3
        hi[axis] = slice(1, None)
        lo[axis] = slice(None, -1)
        out[tuple(lo)] = F[tuple(hi)] - F[tuple(lo)]
    return out


def _curl(F, diff):
    """curl from six one-sided differences d[c][a] = diff(F_c, a).

    component a of the curl is d[w][u] - d[u][w] with (a, u, w) cyclic, exactly
    the `+=` then `-=` pairs of fdtd/grid.py:43-49 and 68-74.
    """
    d = {}
    for c in range(3):
        for a in range(3):
            if a != c:
                d[(c, a)] = diff(F[..., c], a)
    curl = lib.zeros(F.shape)
    for a in range(3):
        u, w = (a + 1) % 3, (a + 2) % 3
        curl[..., a] = d[(w, u)] - d[(u, w)]
    return curl, d


def curl_H(H):
    """E-type curl (backward differences) -- fdtd/grid.py:54-76."""
    return _curl(H, _bdiff)[0]


def curl_E(E):
    """H-type curl (forward differences) -- fdtd/grid.py:29-51."""
    return _curl(E, _fdiff)[0]


# --------------------------------------------------------------------------- grid
class Grid:
    """State and stepping order of fdtd/grid.py:80-331."""

    def __init__(self, shape, grid_spacing=155e-9, permittivity=1.0,
                 permeability=1.0, courant_number=None):
        self.grid_spacing = float(grid_spacing)
        if len(shape) != 3:
            raise ValueError(f"invalid grid shape {shape}")
        self.Nx, self.Ny, self.Nz = (self._cells(s) for s in shape)
        self.D = int(self.Nx > 1) + int(self.Ny > 1) + int(self.Nz > 1)
        cmax = float(self.D) ** (-0.5)
        if courant_number is None:
            self.courant_number = 0.99 * cmax
        elif courant_number > cmax:
            raise ValueError("courant_number too high")
        else:
            self.courant_number = float(courant_number)
        self.time_step = self.courant_number * self.grid_spacing / C0
        full = (self.Nx, self.Ny, self.Nz, 3)
        self.E = lib.zeros(full)
        self.H = lib.zeros(full)
        self.inverse_permittivity = lib.ones(full) / self._material(permittivity)
        self.inverse_permeability = lib.ones(full) / self._material(permeability)
        self.time_steps_passed = 0
        self.sources, self.boundaries, self.detectors, self.objects = [], [], [], []

    @staticmethod
    def _material(value):
        if lib.is_array(value) and len(value.shape) == 3:
            value = value[:, :, :, None]
        return lib.asarray(value)

    # index handling: fdtd/grid.py:171-223
    def _cells(self, d):
        if not isinstance(d, int):
            return int(float(d) / self.grid_spacing + 0.5)
        return d

    def _steps(self, t):
        if not isinstance(t, int):
            return int(float(t) / self.time_step + 0.5)
        return t

    def _key(self, key):
        try:
            len(key)
            return [self._cells(k) for k in key]
        except TypeError:
            if isinstance(key, slice):
                f = lambda v: self._cells(v) if isinstance(v, float) else v
                return slice(f(key.start), f(key.stop), f(key.step))
            return [self._cells(key)]

    def __setitem__(self, key, attr):
        if not isinstance(key, tuple):
            x, y, z = key, slice(None), slice(None)
        elif len(key) == 1:
            x, y, z = key[0], slice(None), slice(None)
        elif len(key) == 2:
            x, y, z = key[0], key[1], slice(None)
        elif len(key) == 3:
            x, y, z = key
        else:
            raise KeyError("maximum number of indices for the grid is 3")
        attr._register_grid(grid=self, x=self._key(x), y=self._key(y), z=self._key(z))

    @property
    def shape(self):
        return (self.Nx, self.Ny, self.Nz)

    @property
    def time_passed(self):
        return self.time_steps_passed * self.time_step

    def _name(self, thing):
        if thing.name is not None:
            if hasattr(self, thing.name):
                raise ValueError(f"The grid already has an attribute with name {thing.name}")
            setattr(self, thing.name, thing)

    # stepping: fdtd/grid.py:250-331
    def run(self, total_time, progress_bar=False):
        if isinstance(total_time, float):
            total_time /= self.time_step
        for _ in range(int(total_time)):
            self.step()

    def step(self):
        self.update_E()
        self.update_H()
        self.time_steps_passed += 1

    def update_E(self):
        sc = self.courant_number
        curl, d = _curl(self.H, _bdiff)
        for b in self.boundaries:
            b.advance_psi_E(d)
        self.E += sc * self.inverse_permittivity * curl
        for o in self.objects:
            o.update_E(curl)
        for b in self.boundaries:
            b.apply_E()
        for s in self.sources:
            s.update_E()
        for det in self.detectors:
            det.detect_E()

    def update_H(self):
        sc = self.courant_number
        curl, d = _curl(self.E, _fdiff)
        for b in self.boundaries:
            b.advance_psi_H(d)
        self.H -= sc * self.inverse_permeability * curl
        for b in self.boundaries:
            b.apply_H()
        for s in self.sources:
            s.update_H()
        for det in self.detectors:
            det.detect_H()

    def reset(self):
        self.H *= 0.0
        self.E *= 0.0
        self.time_steps_passed *= 0


# --------------------------------------------------------------------------- boundaries
class _Boundary:
    def __init__(self, name=None):
        self.grid = None
        self.name = name

    def advance_psi_E(self, d):
        pass

    def advance_psi_H(self, d):
        pass

    def apply_E(self):
        pass

    def apply_H(self):
        pass


class PeriodicBoundary(_Boundary):
    """fdtd/boundaries.py:143-219: E[0] = E[-1] after the E update, H[-1] = H[0] after H."""

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.boundaries.append(self)
        pos = []
        for s in (x, y, z):
            if isinstance(s, list):
                if len(s) > 1:
                    raise ValueError("Use slices or single numbers to index the grid for a boundary")
                pos.append(s[0])
            elif isinstance(s, slice):
                if (s.start is not None and s.stop is not None
                        and (s.start == s.stop or abs(s.start - s.stop) == 1)):
                    pos.append(s.start)
                else:
                    pos.append(s)
            else:
                raise ValueError("Invalid grid indexing used for boundary")
        grid._name(self)
        for axis, p in enumerate(pos):
            if not isinstance(p, slice) and (p == 0 or p == -1):
                lo, hi = (f"_{'xyz'[axis]}low_boundary", f"_{'xyz'[axis]}high_boundary")
                if hasattr(grid, lo) or hasattr(grid, hi):
                    raise AttributeError("grid already has a boundary there!")
                setattr(grid, lo, self)
                setattr(grid, hi, self)
                self.axis = axis
                return
        raise IndexError("A periodic boundary should be placed at the boundary of the grid "
                         "using a single index (either 0 or -1)")

    def _planes(self):
        first = [slice(None)] * 3
        last = [slice(None)] * 3
        first[self.axis] = 0
        last[self.axis] = -1
        return tuple(first), tuple(last)

    def apply_E(self):
        first, last = self._planes()
        self.grid.E[first] = self.grid.E[last]

    def apply_H(self):
        first, last = self._planes()
        self.grid.H[last] = self.grid.H[first]


class PML(_Boundary):
    """CPML slab, fdtd/boundaries.py:225-625, restated with two psi scalars per cell.

    Slab on axis a (u = a+1, w = a+2 cyclic), local index l:
        psiE0 <- psiE0*bE[l] + [l>=1] dH(w,a)*cE[l]     (drives E_u with sign -)
        psiE1 <- psiE1*bE[l] + [l>=1] dH(u,a)*cE[l]     (drives E_w with sign +)
        psiH0 <- psiH0*bH[l] + [l<t-1] dE(w,a)*cH[l]    (drives H_u with sign -)
        psiH1 <- psiH1*bH[l] + [l<t-1] dE(u,a)*cH[l]    (drives H_w with sign +)
    """

    def __init__(self, a=1e-8, name=None):
        super().__init__(name)
        self.k = 1.0
        self.a = a
        self.thickness = 0

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.boundaries.append(self)
        for s in (x, y, z):
            if isinstance(s, list):
                raise ValueError("One can only use slices to index the grid for a PML")
            if not isinstance(s, slice):
                raise ValueError("Invalid grid indexing used for boundary")
        grid._name(self)
        # orientation: first matching axis, low before high (fdtd/boundaries.py:301-358)
        for axis, s in enumerate((x, y, z)):
            nm = "xyz"[axis]
            if (s.start is None or s.start == 0) and s.stop is not None and s.stop > 0:
                side, t = "low", s.stop
            elif s.start is not None and s.stop is None and s.start < 0:
                side, t = "high", -s.start
            else:
                continue
            if hasattr(grid, f"_{nm}{side}_boundary"):
                raise AttributeError(f"grid already has an {nm}{side} boundary!")
            setattr(grid, f"_{nm}{side}_boundary", self)
            self.axis, self.side, self.thickness = axis, side, t
            self._setup()
            return
        raise IndexError("not a valid slice for a PML. Make sure the slice is at the border of the PML")

    def _setup(self):
        g, t, a = self.grid, self.thickness, self.axis
        sc = g.courant_number
        # sigma profiles: fdtd/boundaries.py:291-293 and the _set_sigmaE/_set_sigmaH of each orientation
        sig = lambda v: 40 * v ** 3 / (t + 1) ** 4
        sE = lib.zeros((t,))
        sH = lib.zeros((t,))
        if self.side == "low":
            sE[:] = sig(lib.arange(t - 0.5, -0.5, -1.0))
            sH[:-1] = sig(lib.arange(t - 1.0, 0, -1.0))
        else:
            sE[:] = sig(lib.arange(0.5, t + 0.5, 1.0))
            sH[:-1] = sig(lib.arange(1.0, t, 1.0))
        # fdtd/boundaries.py:396-407
        self.bE = lib.exp(-(sE / self.k + self.a) * sc)
        self.cE = (self.bE - 1.0) * sE / (sE * self.k + self.a * self.k ** 2)
        self.bH = lib.exp(-(sH / self.k + self.a) * sc)
        self.cH = (self.bH - 1.0) * sH / (sH * self.k + self.a * self.k ** 2)
        n = [g.Nx, g.Ny, g.Nz]
        loc = [slice(None)] * 3
        loc[a] = slice(None, t) if self.side == "low" else slice(-t, None)
        self.loc = tuple(loc)
        shp = list(n)
        shp[a] = t
        self.psiE = [lib.zeros(shp), lib.zeros(shp)]
        self.psiH = [lib.zeros(shp), lib.zeros(shp)]
        bshape = [1, 1, 1]
        bshape[a] = t
        self._b = lambda v: v.reshape(bshape)
        inner_hi = [slice(None)] * 3
        inner_hi[a] = slice(1, None)
        inner_lo = [slice(None)] * 3
        inner_lo[a] = slice(None, -1)
        self._hi, self._lo = tuple(inner_hi), tuple(inner_lo)

    def _advance(self, psi, b, c, d, sel):
        a = self.axis
        u, w = (a + 1) % 3, (a + 2) % 3
        b3, c3 = self._b(b), self._b(c)
        for n, comp in enumerate((w, u)):
            psi[n] *= b3
            diff = d[(comp, a)][self.loc]
            psi[n][sel] += diff[sel] * c3[sel]

    def advance_psi_E(self, d):
        self._advance(self.psiE, self.bE, self.cE, d, self._hi)

    def advance_psi_H(self, d):
        self._advance(self.psiH, self.bH, self.cH, d, self._lo)

    def _apply(self, F, inv, psi, sign):
        a = self.axis
        u, w = (a + 1) % 3, (a + 2) % 3
        sc = self.grid.courant_number
        Fs, invs = F[self.loc], inv[self.loc]
        phi_u = 0.0 - psi[0]
        phi_w = psi[1] - 0.0
        if sign > 0:
            Fs[..., u] += sc * invs[..., u] * phi_u
            Fs[..., w] += sc * invs[..., w] * phi_w
        else:
            Fs[..., u] -= sc * invs[..., u] * phi_u
            Fs[..., w] -= sc * invs[..., w] * phi_w

    def apply_E(self):
        self._apply(self.grid.E, self.grid.inverse_permittivity, self.psiE, +1)

    def apply_H(self):
        self._apply(self.grid.H, self.grid.inverse_permeability, self.psiH, -1)


# --------------------------------------------------------------------------- objects
class Object:
    """fdtd/objects.py:24-129."""

    def __init__(self, permittivity, name=None):
        self.grid = None
        self.name = name
        self.permittivity = lib.asarray(permittivity)

    @staticmethod
    def _norm(s, n):
        if isinstance(s, list):
            if len(s) == 1:
                return slice(s[0], s[0] + 1, None)
            raise IndexError("One can only use slices or single indices to index the grid for an Object")
        if isinstance(s, slice):
            start, stop, step = s.start, s.stop, s.step
            if step is not None and step != 1:
                raise IndexError("Can only use slices with unit step to index the grid for an Object")
            start = 0 if start is None else start
            start = n + start if start < 0 else start
            stop = n if stop is None else stop
            stop = n + stop if stop < 0 else stop
            return slice(start, stop, None)
        raise ValueError("Invalid grid indexing used for object")

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.objects.append(self)
        grid._name(self)
        self.x, self.y, self.z = (self._norm(x, grid.Nx), self._norm(y, grid.Ny),
                                  self._norm(z, grid.Nz))
        self.Nx = abs(self.x.stop - self.x.start)
        self.Ny = abs(self.y.stop - self.y.start)
        self.Nz = abs(self.z.stop - self.z.start)
        eps = self.permittivity
        if lib.is_array(eps) and len(eps.shape) == 3:
            eps = eps[:, :, :, None]
        self.inverse_permittivity = lib.ones((self.Nx, self.Ny, self.Nz, 3)) / eps
        gi = grid.inverse_permittivity
        # border fix uses the GRID's last plane (fdtd/objects.py:79-90)
        if self.Nx > 1:
            self.inverse_permittivity[-1, :, :, 0] = gi[-1, self.y, self.z, 0]
        if self.Ny > 1:
            self.inverse_permittivity[:, -1, :, 1] = gi[self.x, -1, self.z, 1]
        if self.Nz > 1:
            self.inverse_permittivity[:, :, -1, 2] = gi[self.x, self.y, -1, 2]
        gi[self.x, self.y, self.z] = 0

    def update_E(self, curl):
        loc = (self.x, self.y, self.z)
        g = self.grid
        g.E[loc] = g.E[loc] + g.courant_number * self.inverse_permittivity * curl[loc]


class AbsorbingObject(Object):
    """fdtd/objects.py:163-221."""

    def __init__(self, permittivity, conductivity, name=None):
        super().__init__(permittivity, name)
        self.conductivity = lib.asarray(conductivity)

    def _register_grid(self, grid, x, y, z):
        super()._register_grid(grid, x, y, z)
        s = self.conductivity
        while s.ndim < 4:
            s = s[..., None]
        self.absorption_factor = (0.5 * grid.courant_number * self.inverse_permittivity
                                  * s * grid.grid_spacing * ETA0)

    def update_E(self, curl):
        loc = (self.x, self.y, self.z)
        g, f = self.grid, self.absorption_factor
        g.E[loc] *= (1 - f) / (1 + f)
        g.E[loc] += g.courant_number * self.inverse_permittivity * curl[loc] / (1 + f)


class AnisotropicObject(Object):
    """fdtd/objects.py:232-269: E += sc * (diag(eps^-1) @ curl); the matrices are diagonal."""

    def update_E(self, curl):
        loc = (self.x, self.y, self.z)
        g = self.grid
        g.E[loc] += g.courant_number * (self.inverse_permittivity * curl[loc])


# --------------------------------------------------------------------------- sources
def hanning(f, t, n):
    """fdtd/waveforms.py:8-9."""
    return (1 / 2) * (1 - cos(f * t / n)) * (sin(f * t))


def _line_points(grid, x, y, z, convert, min_points):
    """Diagonal point list of Line sources / detectors (fdtd/sources.py:209-276,
    fdtd/detectors.py:62-112).  `convert` applies the metre->cell conversion a
    second time, as LineSource does and LineDetector does not."""
    c = grid._cells if convert else (lambda v: v)
    if isinstance(x, list) and isinstance(y, list) and isinstance(z, list):
        if len(x) != len(y) or len(y) != len(z):
            raise IndexError("sources require grid to be indexed with slices or equal length list-indices")
        return [c(v) for v in x], [c(v) for v in y], [c(v) for v in z]
    sl = []
    for s, n in ((x, grid.Nx), (y, grid.Ny), (z, grid.Nz)):
        if isinstance(s, list):
            s = slice(c(s[0]), c(s[-1]), None)
        a = c(s.start if s.start is not None else 0)
        b = c(s.stop if s.stop is not None else n)
        sl.append((a, b))
    m = max(abs(b - a) for a, b in sl)
    if m < min_points:
        raise ValueError("a LineSource should consist of at least two gridpoints")
    pts = [[int(v) for v in _np.linspace(a, b, m, endpoint=False).astype(_np.int64)]
           for a, b in sl]
    return pts[0], pts[1], pts[2]


class _Waveform:
    def _wave(self, q):
        """per-step scalar of Point/Line sources (fdtd/sources.py:93-108, 278-295)."""
        if self.pulse:
            t1 = int(2 * pi / (self.frequency * self.hanning_dt / self.cycle))
            if q < t1:
                return hanning(self.frequency, q * self.hanning_dt, self.cycle)
            return 0
        return sin(2 * pi * q / self.period + self.phase_shift)


class PointSource(_Waveform):
    """fdtd/sources.py:25-127 -- soft source on Ez."""

    def __init__(self, period=15, amplitude=1.0, phase_shift=0.0, name=None,
                 pulse=False, cycle=5, hanning_dt=10.0):
        self.grid = None
        self.period, self.amplitude, self.phase_shift = period, amplitude, phase_shift
        self.name, self.pulse, self.cycle = name, pulse, cycle
        self.frequency = 1.0 / period
        self.hanning_dt = hanning_dt if hanning_dt is not None else 0.5 / self.frequency

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.sources.append(self)
        grid._name(self)
        try:
            (x,), (y,), (z,) = x, y, z
        except (TypeError, ValueError):
            raise ValueError("a point source should be placed on a single grid cell.")
        self.x, self.y, self.z = grid._cells(x), grid._cells(y), grid._cells(z)
        self.period = grid._steps(self.period)
        self.frequency = 1.0 / self.period

    def update_E(self):
        q = self.grid.time_steps_passed
        self.grid.E[self.x, self.y, self.z, 2] += self.amplitude * self._wave(q)

    def update_H(self):
        pass


class LineSource(_Waveform):
    """fdtd/sources.py:131-315 -- soft source on Ez along the box diagonal."""

    def __init__(self, period=15, amplitude=1.0, phase_shift=0.0, name=None,
                 pulse=False, cycle=5, hanning_dt=10.0):
        self.grid = None
        self.period, self.amplitude, self.phase_shift = period, amplitude, phase_shift
        self.name, self.pulse, self.cycle = name, pulse, cycle
        self.frequency = 1.0 / period
        self.hanning_dt = hanning_dt if hanning_dt is not None else 0.5 / self.frequency

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.sources.append(self)
        grid._name(self)
        self.x, self.y, self.z = _line_points(grid, x, y, z, convert=True, min_points=2)
        self.period = grid._steps(self.period)
        self.frequency = 1.0 / self.period
        L = len(self.x)
        ix, iy, iz = (_np.array(v) for v in (self.x, self.y, self.z))
        vect = lib.asarray((ix - self.x[L // 2]) ** 2 + (iy - self.y[L // 2]) ** 2
                           + (iz - self.z[L // 2]) ** 2)
        self.profile = lib.exp(-(vect ** 2) / (2 * (0.5 * vect.max()) ** 2))
        self.profile /= self.profile.sum()
        self.profile *= self.amplitude

    def update_E(self):
        q = self.grid.time_steps_passed
        vect = self.profile * self._wave(q)
        E = self.grid.E
        for x, y, z, value in zip(self.x, self.y, self.z, vect):
            E[x, y, z, 2] += value

    def update_H(self):
        pass


class PlaneSource:
    """fdtd/sources.py:319-501 -- hard source on one E and one H component."""

    def __init__(self, period=15, amplitude=1.0, phase_shift=0.0, name=None, polarization="z"):
        self.grid = None
        self.period, self.amplitude, self.phase_shift = period, amplitude, phase_shift
        self.name, self.polarization = name, polarization

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.sources.append(self)
        grid._name(self)
        sl = []
        for s, n in ((x, grid.Nx), (y, grid.Ny), (z, grid.Nz)):
            if not isinstance(s, slice):
                if isinstance(s, list):
                    (s,) = s
                s = slice(grid._cells(s), grid._cells(s) + 1, None)
            a = grid._cells(s.start if s.start is not None else 0)
            b = grid._cells(s.stop if s.stop is not None else n)
            sl.append(slice(a, b) if a < b else (slice(b, a) if a > b else slice(a, a + 1)))
        ext = [s.stop - s.start for s in sl]
        if ext.count(0) > 0:
            raise ValueError("Given location for PlaneSource results in slices of length 0!")
        if ext.count(1) == 0:
            raise ValueError("Given location for PlaneSource is not a 2D plane!")
        if ext.count(1) > 1:
            raise ValueError("Given location for PlaneSource should have no more than one "
                             "dimension in which it's flat.")
        self._Epol = "xyz".index(self.polarization)
        if ext[self._Epol] == 1:
            raise ValueError("PlaneSource cannot be polarized perpendicular to the orientation of the plane.")
        # fdtd/sources.py:468-472
        probe, first, second = [(2, 1, 2), (2, 0, 2), (1, 0, 1)][self._Epol]
        self._Hpol = first if ext[probe] == 1 else second
        self.x, self.y, self.z = sl
        self.period = grid._steps(self.period)
        self.frequency = 1.0 / self.period
        self.profile = self.amplitude * lib.ones(ext)

    def _value(self):
        q = self.grid.time_steps_passed
        return self.profile * sin(2 * pi * q / self.period + self.phase_shift)

    def update_E(self):
        self.grid.E[self.x, self.y, self.z, self._Epol] = self._value()

    def update_H(self):
        self.grid.H[self.x, self.y, self.z, self._Hpol] = self._value()


# --------------------------------------------------------------------------- detectors
class LineDetector:
    """fdtd/detectors.py:20-139: one (L,3) sample of E and of H per step."""

    def __init__(self, name=None):
        self.grid = None
        self.E, self.H = [], []
        self.name = name

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.detectors.append(self)
        grid._name(self)
        self.x, self.y, self.z = _line_points(grid, x, y, z, convert=False, min_points=0)

    def detect_E(self):
        self.E.append(lib.copy(self.grid.E[self.x, self.y, self.z]))

    def detect_H(self):
        self.H.append(lib.copy(self.grid.H[self.x, self.y, self.z]))

    def detector_values(self):
        return {"E": self.E, "H": self.H}


class BlockDetector(LineDetector):
    """fdtd/detectors.py:146-280: inclusive ranges, nested [i][j][k] -> (3,) samples.

    The oracle stores each step as one (nx,ny,nz,3) array; nesting it into lists
    does not change any value."""

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.detectors.append(self)
        grid._name(self)
        if isinstance(x, list) and isinstance(y, list) and isinstance(z, list):
            if len(x) != len(y) or len(y) != len(z):
                raise IndexError("sources require grid to be indexed with slices or equal length list-indices")
            self.x, self.y, self.z = x, y, z
            return
        out = []
        for s, n in ((x, grid.Nx), (y, grid.Ny), (z, grid.Nz)):
            if isinstance(s, list):
                s = slice(s[0], s[-1], None)
            a = s.start if s.start is not None else 0
            b = s.stop if s.stop is not None else n
            out.append(list(range(a, b + 1)))
        self.x, self.y, self.z = out

    def _sample(self, F):
        ix = _np.array(self.x)[:, None, None]
        iy = _np.array(self.y)[None, :, None]
        iz = _np.array(self.z)[None, None, :]
        return lib.copy(F[ix, iy, iz])

    def detect_E(self):
        self.E.append(self._sample(self.grid.E))

    def detect_H(self):
        self.H.append(self._sample(self.grid.H))


class CurrentDetector(BlockDetector):
    """fdtd/detectors.py:284-496: z-current from the H loop around a cell, two z levels averaged.
    One (nx,ny,nz) array per H half-step in `I`; nothing is recorded after the E half-step."""

    def __init__(self, name=None):
        super().__init__(name)
        self.I = []

    def detect_E(self):
        pass

    def _point(self, px, py, pz):
        H, dx = self.grid.H, self.grid.grid_spacing
        # fdtd/detectors.py:444-461 -- note the `+=` on the second current_vector_2
        v1 = (H[px, py - 1, pz, 0] - H[px, py, pz, 0]) * dx
        v2 = (H[px, py, pz, 1] - H[px - 1, py, pz, 1]) * dx
        c1 = v1 + v2
        v1 = (H[px, py - 1, pz - 1, 0] - H[px, py, pz - 1, 0]) * dx
        v2 = v2 + (H[px, py, pz - 1, 1] - H[px - 1, py, pz - 1, 1]) * dx
        c2 = v1 + v2
        return (c1 + c2) / 2.0

    def detect_H(self):
        out = lib.zeros((len(self.x), len(self.y), len(self.z)))
        for i, px in enumerate(self.x):
            for j, py in enumerate(self.y):
                for k, pz in enumerate(self.z):
                    out[i, j, k] = self._point(px, py, pz)
        self.I.append(out)

    def detector_values(self):
        return {"I": self.I}


class SoftArbitraryPointSource:
    """fdtd/sources.py:504-643: Ez += (waveform[q] + Z*I_prev)/dx with I_prev the paired
    CurrentDetector's sample of the previous step."""

    def __init__(self, waveform_array, name=None, impedance=0.0):
        self.grid = None
        self.name = name
        self.waveform_array = waveform_array
        self.impedance = impedance
        self.input_voltage, self.source_voltage = [], []

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        grid.sources.append(self)
        grid._name(self)
        try:
            (x,), (y,), (z,) = x, y, z
        except (TypeError, ValueError):
            raise ValueError("a point source should be placed on a single grid cell.")
        self.x, self.y, self.z = grid._cells(x), grid._cells(y), grid._cells(z)
        self.current_detector = CurrentDetector(name=None if self.name is None else self.name + "_I")
        grid[x, y, z] = self.current_detector

    def update_E(self):
        q = self.grid.time_steps_passed
        vin = self.waveform_array[q] if q < self.waveform_array.shape[0] else 0.0
        cur = self.current_detector.I[-1][0][0][0] if q > 0 else 0.0
        vout = vin + self.impedance * cur if self.impedance > 0 else vin
        self.grid.E[self.x, self.y, self.z, 2] += vout / self.grid.grid_spacing
        self.input_voltage.append([[[vin]]])
        self.source_voltage.append([[[vout]]])

    def update_H(self):
        pass
