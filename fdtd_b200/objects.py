"""Objects: regions of the grid with their own permittivity (fdtd/objects.py).

Registration reproduces the reference exactly -- slice normalisation (fdtd/objects.py:94-116),
`ones/permittivity`, the border fix that copies the GRID's last plane into the object's last
plane (:79-90), zeroing the grid's eps^-1 inside the region (:92), the absorption factor
(:198-205).  The per-step `update_E` of every object (:118-129, :207-221, :254-269) is not a
Python call here: the engine folds all objects into per-cell coefficient arrays that the fused
E half-step kernel streams only in the tiles an object touches (hundreds of objects, as in the
reference's lens examples, cost nothing per step).  Objects may overlap to any depth: the first two
covering a cell are the kernel's two coefficient layers, further ones run as small kernels of their own.
"""
import torch
import torch.distributed as dist

from . import constants as const
from .backend import backend as bd


class Object:
    """An object to place in the grid."""

    def __init__(self, permittivity, name: str = None):
        self.grid = None
        self.name = name
        if bd.is_complex(permittivity):
            raise NotImplementedError("complex permittivity is not supported by the CUDA engine")
        self.permittivity = bd.require().array(permittivity)

    def _register_grid(self, grid, x, y, z):
        self.x = self._handle_slice(x, max_index=grid.Nx)
        self.y = self._handle_slice(y, max_index=grid.Ny)
        self.z = self._handle_slice(z, max_index=grid.Nz)
        self.Nx = abs(self.x.stop - self.x.start)
        self.Ny = abs(self.y.stop - self.y.start)
        self.Nz = abs(self.z.stop - self.z.start)
        self.grid = grid
        self.grid.objects.append(self)
        grid._register_name(self)

        part = grid._part
        lx0, lx1 = part.local_range(self.x.start, self.x.stop)
        self._nx_local = lx1 - lx0
        self._loc = (slice(lx0, lx1), self.y, self.z)          # into (nx_local, Ny, Nz) arrays
        ox0 = lx0 + part.x0 - self.x.start                      # first object plane held locally

        eps = self.permittivity
        if eps.dim() == 3:
            eps = eps[:, :, :, None]
        if eps.dim() == 4 and eps.shape[0] == self.Nx and self.Nx != 1:
            eps = eps[ox0:ox0 + self._nx_local]                 # local x-part of a per-cell array
        inv = bd.ones((self._nx_local, self.Ny, self.Nz, 3), dtype=self.permittivity.dtype) / eps

        gi = grid._materialize("eps")                           # (3, nx_local, Ny, Nz)
        # border fix (fdtd/objects.py:79-90): the object's last plane along each axis takes the
        # GRID's last plane (global index -1) at registration time
        if self.Nx > 1:
            last = self._grid_last_x_plane(gi)                  # (Ny_obj, Nz_obj), on every rank
            if self._nx_local > 0 and self.x.stop - 1 < part.x1 and self.x.stop - 1 >= part.x0:
                inv[-1, :, :, 0] = last
        if self.Ny > 1 and self._nx_local > 0:
            inv[:, -1, :, 1] = gi[1, lx0:lx1, -1, self.z]
        if self.Nz > 1 and self._nx_local > 0:
            inv[:, :, -1, 2] = gi[2, lx0:lx1, self.y, -1]
        self.inverse_permittivity = inv                          # (nx_local, Ny, Nz, 3), reference layout
        self._inv_eps_soa = inv.permute(3, 0, 1, 2).contiguous() # (3, nx_local, Ny, Nz) for the bake
        self._absorb_soa = None
        if self._nx_local > 0:
            # zero inside the object (fdtd/objects.py:92); the bake later turns the zeros of AnisotropicObject
            # cells into NEGATIVE zeros as a marker for the kernel (include/fdtd_b200.h, FDTD_CLS_ANISO)
            gi[:, lx0:lx1, self.y, self.z] = 0.0

    def _grid_last_x_plane(self, gi):
        """grid.inverse_permittivity[-1, y, z, 0] -- lives on the last rank when sharded."""
        part = self.grid._part
        if not part.sharded:
            return gi[0, -1, self.y, self.z].clone()
        buf = gi[0, -1, self.y, self.z].clone().contiguous()
        dist.broadcast(buf, src=part.world - 1)
        return buf

    # the reference's plug-in protocol (fdtd/grid.py:285-287, 311-313).  The engine folds the built-in object kinds
    # into the fused kernel and never calls these; a subclass that OVERRIDES them is driven through the per-step
    # plug-in path instead (fdtd_b200/engine.py, `_hooks`)
    def update_E(self, curl_H):
        pass

    def update_H(self, curl_E):
        pass

    update_E._fdtd_b200_builtin = update_H._fdtd_b200_builtin = True

    def _handle_slice(self, s, max_index: int = None) -> slice:
        if isinstance(s, list):
            if len(s) == 1:
                return slice(s[0], s[0] + 1, None)
            raise IndexError("One can only use slices or single indices to index the grid for an Object")
        if isinstance(s, slice):
            start, stop, step = s.start, s.stop, s.step
            if step is not None and step != 1:
                raise IndexError("Can only use slices with unit step to index the grid for an Object")
            if start is None:
                start = 0
            if start < 0:
                start = max_index + start
            if stop is None:
                stop = max_index
            if stop < 0:
                stop = max_index + stop
            return slice(start, stop, None)
        raise ValueError("Invalid grid indexing used for object")

    def __repr__(self):
        return f"{self.__class__.__name__}(name={repr(self.name)})"

    def __str__(self):
        s = "    " + repr(self) + "\n"

        def fmt(v):
            return str(v).replace("slice(", "").replace(")", "").replace(", ", ":").replace("None", "")

        s += f"        @ x={fmt(self.x)}, y={fmt(self.y)}, z={fmt(self.z)}".replace(":,", ",")
        if s[-1] == ":":
            s = s[:-1]
        return s + "\n"


class AbsorbingObject(Object):
    """An absorbing object takes conductivity into account (fdtd/objects.py:163-229)."""

    def __init__(self, permittivity, conductivity, name: str = None):
        super().__init__(permittivity, name)
        self.conductivity = bd.array(conductivity)

    def _register_grid(self, grid, x=None, y=None, z=None):
        super()._register_grid(grid=grid, x=x, y=y, z=z)
        conductivity = self.conductivity
        while conductivity.dim() < self.inverse_permittivity.dim():
            conductivity = conductivity[..., None]
        if conductivity.shape[0] == self.Nx and self.Nx != 1 and self._nx_local != self.Nx:
            part = grid._part
            o = self._loc[0].start + part.x0 - self.x.start
            conductivity = conductivity[o:o + self._nx_local]
        self.conductivity = torch.broadcast_to(conductivity, self.inverse_permittivity.shape)
        # fdtd/objects.py:198-205, evaluated left to right
        self.absorption_factor = (
            0.5
            * grid.courant_number
            * self.inverse_permittivity
            * self.conductivity
            * grid.grid_spacing
            * const.eta0
        )
        self._absorb_soa = self.absorption_factor.permute(3, 0, 1, 2).contiguous()


class AnisotropicObject(Object):
    """An object with a (diagonal) anisotropic permittivity tensor (fdtd/objects.py:232-277).

    The reference expands eps^-1 to diagonal 3x3 matrices and multiplies with bmm; the
    off-diagonals are exactly zero, so the update is the per-component product the fused kernel
    already performs.  `inverse_permittivity` keeps the reference's (Nx,Ny,Nz,3,3) shape."""

    def _register_grid(self, grid, x=None, y=None, z=None):
        super()._register_grid(grid=grid, x=x, y=y, z=z)
        self.inverse_permittivity = torch.diag_embed(self.inverse_permittivity)
