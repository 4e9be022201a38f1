"""Boundaries: PeriodicBoundary and the CPML, registration semantics of fdtd/boundaries.py.

Only registration and coefficient construction happen here (once); the per-step work --
PML.update_phi_E/H, PML.update_E/H, periodic copies (fdtd/boundaries.py:184-219, 409-487) --
runs inside the fused CUDA kernels.  A PML keeps two psi scalars per slab cell and per field
and four 1-D tables of length `thickness` instead of the reference's 14 slab-sized arrays
(SURVEY.md section 8a rows P1-P5).
"""
from .backend import backend as bd
from ._hostmath import HostLib


class Boundary:
    """base class: name handling and the slice checks of fdtd/boundaries.py:19-77."""

    def __init__(self, name: str = None):
        self.grid = None
        self.name = name

    def _register_grid(self, grid, x, y, z):
        self.grid = grid
        self.grid.boundaries.append(self)
        self.x = self._handle_slice(x)
        self.y = self._handle_slice(y)
        self.z = self._handle_slice(z)
        grid._register_name(self)

    # the reference's plug-in protocol (fdtd/grid.py:279-281, 290-291, 305-307, 316-317): done inside the kernels for
    # the built-in boundaries; overridden in a subclass, they are called per step (fdtd_b200/engine.py, `_hooks`)
    def update_phi_E(self):
        pass

    def update_phi_H(self):
        pass

    def update_E(self):
        pass

    def update_H(self):
        pass

    for _m in (update_phi_E, update_phi_H, update_E, update_H):
        _m._fdtd_b200_builtin = True
    del _m

    def _handle_slice(self, s):
        if isinstance(s, list):
            if len(s) > 1:
                raise ValueError("Use slices or single numbers to index the grid for a boundary")
            return s[0]
        if isinstance(s, slice):
            if (s.start is not None and s.stop is not None
                    and (s.start == s.stop or abs(s.start - s.stop) == 1)):
                return s.start
            return s
        raise ValueError("Invalid grid indexing used for boundary")

    def __repr__(self):
        return f"PML(name={repr(self.name)})"   # sic: the reference prints every boundary as PML

    def __str__(self):
        s = "    " + repr(self) + "\n"

        def fmt(v):
            return str(v).replace("slice(", "").replace(")", "").replace(", ", ":").replace("None", "")

        s += f"        @ x={fmt(self.x)}, y={fmt(self.y)}, z={fmt(self.z)}".replace(":,", ",")
        if s[-1] == ":":
            s = s[:-1]
        return s + "\n"


class PeriodicBoundary(Boundary):
    """E[0] = E[-1] after every E half-step and H[-1] = H[0] after every H half-step on one
    axis (fdtd/boundaries.py:143-219), executed in registration order among the boundaries."""

    def _register_grid(self, grid, x, y, z):
        super()._register_grid(grid=grid, x=x, y=y, z=z)
        for axis, pos in enumerate((self.x, self.y, self.z)):
            if isinstance(pos, slice) or not (pos == 0 or pos == -1):
                continue
            nm = "xyz"[axis]
            if hasattr(grid, f"_{nm}low_boundary") or hasattr(grid, f"_{nm}high_boundary"):
                raise AttributeError(f"grid already has an {nm}low/{nm}high boundary!")
            setattr(grid, f"_{nm}low_boundary", self)
            setattr(grid, f"_{nm}high_boundary", self)
            self.axis = axis
            return
        raise IndexError("A periodic boundary should be placed at the boundary of the "
                         "grid using a single index (either 0 or -1)")


class PML(Boundary):
    """Convolutional perfectly matched layer (fdtd/boundaries.py:225-625)."""

    def __init__(self, a: float = 1e-8, name: str = None):
        super().__init__(name=name)
        self.k = 1.0
        self.thickness = 0
        self.a = a
        if a == 0:
            # the reference divides 0 by 0 where sigma = 0 (fdtd/boundaries.py:396-400) and fills
            # the grid with NaN; refuse instead (SURVEY.md 8a trap 10)
            raise ValueError("PML stability parameter a must be non-zero")

    def _handle_slice(self, s):
        if isinstance(s, list):
            raise ValueError("One can only use slices to index the grid for a PML")
        if isinstance(s, slice):
            return s
        raise ValueError("Invalid grid indexing used for boundary")

    def _register_grid(self, grid, x, y, z):
        super()._register_grid(grid=grid, x=x, y=y, z=z)
        # orientation: first axis whose slice touches a face, low side tested before high side
        # (fdtd/boundaries.py:301-358)
        for axis, s in enumerate((self.x, self.y, self.z)):
            nm = "xyz"[axis]
            if (s.start is None or s.start == 0) and (s.stop is not None) and (s.stop > 0):
                side, t = "low", s.stop
            elif (s.start is not None) and (s.stop is None) and (s.start < 0):
                side, t = "high", -s.start
            else:
                continue
            if hasattr(grid, f"_{nm}{side}_boundary"):
                raise AttributeError(f"grid already has an {nm}{side} boundary!")
            setattr(grid, f"_{nm}{side}_boundary", self)
            self.axis, self.side = axis, side
            self._calculate_parameters(thickness=t)
            return
        raise IndexError("not a valid slice for a PML. Make sure the slice is at the border of the PML")

    def _sigma(self, vect):
        """cubic conductivity profile (fdtd/boundaries.py:291-293)."""
        return 40 * vect ** 3 / (self.thickness + 1) ** 4

    def _calculate_parameters(self, thickness: int = 10):
        g = self.grid
        t = self.thickness = thickness
        n_axis = (g.Nx, g.Ny, g.Nz)[self.axis]
        if t > n_axis:
            raise IndexError(f"PML thickness {t} exceeds the grid extent {n_axis}")
        self.lo = 0 if self.side == "low" else n_axis - t
        hl = HostLib(g._dtype)
        sE, sH = hl.zeros(t), hl.zeros(t)
        # sigma at the E and H sample positions of each orientation (fdtd/boundaries.py:502-625)
        if self.side == "low":
            sE[:] = self._sigma(hl.arange(t - 0.5, -0.5, -1.0))
            sH[:-1] = self._sigma(hl.arange(t - 1.0, 0, -1.0))
        else:
            sE[:] = self._sigma(hl.arange(0.5, t + 0.5, 1.0))
            sH[:-1] = self._sigma(hl.arange(1.0, t, 1.0))
        sc = g.courant_number
        # fdtd/boundaries.py:396-407
        bE = hl.exp(-(sE / self.k + self.a) * sc)
        cE = (bE - 1.0) * sE / (sE * self.k + self.a * self.k ** 2)
        bH = hl.exp(-(sH / self.k + self.a) * sc)
        cH = (bH - 1.0) * sH / (sH * self.k + self.a * self.k ** 2)
        self.sigmaE_profile, self.sigmaH_profile = sE, sH
        self._tab = {k: hl.to_device(v, bd.device) for k, v in (("bE", bE), ("cE", cE), ("bH", bH), ("cH", cH))}
        # psi storage for the local part of the slab (include/fdtd_b200.h, fdtd_slab)
        part = g._part
        if self.axis == 0:
            self._x0, self._x1 = part.local_range(self.lo, self.lo + t)
            cells = (self._x1 - self._x0) * g.Ny * g.Nz
        elif self.axis == 1:
            self._x0, self._x1 = 0, part.nx
            cells = part.nx * t * g.Nz
        else:
            # z slabs: padded rows so that the kernels' 128-bit accesses stay aligned (fdtd_b200.h)
            self._x0, self._x1 = 0, part.nx
            self._row_lo = self.lo & ~3
            self._row = ((self.lo + t - self._row_lo) + 3) & ~3
            cells = part.nx * g.Ny * self._row
        self._psi_E = bd.zeros((2, cells), dtype=g._sdtype)
        self._psi_H = bd.zeros((2, cells), dtype=g._sdtype)

    # user-visible location, as the reference's `loc` (fdtd/boundaries.py:493-497 etc.)
    @property
    def loc(self):
        loc = [slice(None)] * 4
        loc[self.axis] = slice(None, self.thickness) if self.side == "low" else slice(-self.thickness, None)
        return tuple(loc)

    def psi(self, field="E"):
        """the two live psi scalars, shaped like the slab: (2, nx, ny, nz) (local part when sharded)."""
        g, t = self.grid, self.thickness
        p = self._psi_E if field == "E" else self._psi_H
        if self.axis == 0:
            return p.view(2, self._x1 - self._x0, g.Ny, g.Nz)
        if self.axis == 1:
            return p.view(2, g._part.nx, t, g.Nz)
        a = self.lo - self._row_lo
        return p.view(2, g._part.nx, g.Ny, self._row)[..., a:a + t]


def DomainBorderPML(grid, border_cells=5):
    """PML on all six faces (fdtd/boundaries.py:628-656).  Alters the grid in place."""
    if grid.Nx < border_cells * 2 or grid.Ny < border_cells * 2 or grid.Nz < border_cells * 2:
        raise IndexError("PML border_cells larger than domain!")
    grid[:, :, 0:border_cells] = PML()
    grid[:, :, -border_cells:] = PML()
    grid[0:border_cells, :, border_cells:-border_cells] = PML()
    grid[-border_cells:, :, border_cells:-border_cells] = PML()
    grid[border_cells:-border_cells, 0:border_cells, border_cells:-border_cells] = PML()
    grid[border_cells:-border_cells, -border_cells:, border_cells:-border_cells] = PML()
