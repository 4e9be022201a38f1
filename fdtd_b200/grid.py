"""The FDTD Grid -- same public surface as the reference's `fdtd.Grid` (fdtd/grid.py:80-405),
with the per-timestep work done by fused sm_100a CUDA kernels (fdtd_b200/csrc).

Differences a user can see, all deliberate:
  * fields are stored SoA, `(3, Nx+2, Ny, Nz)` with one ghost x-plane at each end;
    `grid.E` / `grid.H` are `(Nx, Ny, Nz, 3)` strided VIEWS of that storage (write-through);
  * `grid.inverse_permittivity` / `inverse_permeability` are created lazily: a homogeneous
    grid never allocates them (they are 12 GiB each at 1024^3 float32) and the kernels then
    stream no coefficient array at all;
  * under an initialised torch.distributed job the grid is split into x-slabs, one per rank
    (`shard="auto"`); `grid.E` then gathers (a copy), `grid.E_local` is the local view.  `x_plane_cost`
    (one relative cost per x-plane) balances the slabs when some planes are dearer, e.g. inside x-PMLs.
"""
import gc
import os

import numpy as np
import torch

from . import constants as const
from .backend import backend as bd
from .sharding import Partition, all_gather_slabs


def _env_flag(name):
    v = os.environ.get(name)
    return None if v in (None, "") else v not in ("0", "false", "False")


def _array_in_array_out(fn):
    """numpy in -> numpy out (the reference's users call the curls on numpy arrays), tensors stay tensors."""
    def wrapped(F):
        if torch.is_tensor(F):
            return fn(F)
        return fn(torch.as_tensor(np.asarray(F))).numpy()
    wrapped.__doc__, wrapped.__name__ = fn.__doc__, fn.__name__
    return wrapped


def _one_sided(F, comp, axis, forward):
    """difference of component `comp` along `axis` towards the next (forward) or from the previous (backward) cell,
    zero on the face where that neighbour does not exist"""
    out = torch.zeros_like(F[..., comp])
    if F.shape[axis] > 1:
        keep = [slice(None)] * 3
        keep[axis] = slice(None, -1) if forward else slice(1, None)
        out[tuple(keep)] = torch.diff(F[..., comp], dim=axis)
    return out


def _curl(F, forward):
    # component a of the curl is dF_w/du - dF_u/dw with (a, u, w) cyclic
    return torch.stack([_one_sided(F, (a + 2) % 3, (a + 1) % 3, forward) - _one_sided(F, (a + 1) % 3, (a + 2) % 3, forward)
                        for a in range(3)], dim=-1)


@_array_in_array_out
def curl_E(E):
    """H-type curl of an (Nx,Ny,Nz,3) array: forward differences, masked on the high faces (fdtd/grid.py:29-51).
    Convenience for user code, the plug-in path and tests; the stepping kernels fuse this and never call it."""
    return _curl(E, forward=True)


@_array_in_array_out
def curl_H(H):
    """E-type curl: backward differences, masked on the low faces (fdtd/grid.py:54-76)."""
    return _curl(H, forward=False)


class Grid:
    """The FDTD grid: owns E, H and the material arrays, and steps them (fdtd/grid.py:80-331)."""

    from .visualization import visualize       # Grid.visualize(x=|y=|z=, ...), fdtd/grid.py:88

    def __init__(self, shape, grid_spacing: float = 155e-9, permittivity=1.0, permeability=1.0,
                 courant_number: float = None, shard="auto", x_plane_cost=None):
        bd.require()
        self.grid_spacing = float(grid_spacing)
        self.Nx, self.Ny, self.Nz = self._handle_tuple(shape)
        self.D = int(self.Nx > 1) + int(self.Ny > 1) + int(self.Nz > 1)
        max_courant_number = float(self.D) ** (-0.5)
        if courant_number is None:
            self.courant_number = 0.99 * max_courant_number
        elif courant_number > max_courant_number:
            raise ValueError(f"courant_number {courant_number} too high for a {self.D}D simulation")
        else:
            self.courant_number = float(courant_number)
        self.time_step = self.courant_number * self.grid_spacing / const.c

        self._dtype = bd.float            # arithmetic and coefficients
        self._sdtype = bd.storage         # the state: E, H, psi, rings ("cuda.float32x": float32 under float64)
        gc.collect()      # grids hold reference cycles (plug-ins point back at them): free dead ones' HBM first
        self._part = Partition(self.Nx, shard, x_plane_cost)
        nx = self._part.nx
        self._E = bd.zeros((3, nx + 2, self.Ny, self.Nz), dtype=self._sdtype)
        self._H = bd.zeros((3, nx + 2, self.Ny, self.Nz), dtype=self._sdtype)
        self._bg_inv_eps, self._inv_eps = self._material(permittivity, "permittivity")
        self._bg_inv_mu, self._inv_mu = self._material(permeability, "permeability")

        self.time_steps_passed = 0
        self.sources, self.boundaries, self.detectors, self.objects = [], [], [], []
        self.folder = None

        self._registration_count = 0
        self._baked_counts = -1
        self._engine = None
        self._ring_fill = {"E": 0, "H": 0}
        self._x_chunk = 0
        # None: automatic (launch-bound small grids), True / False: force.  FDTD_B200_GRAPHS / FDTD_B200_FUSE = 0|1
        self._use_graphs = _env_flag("FDTD_B200_GRAPHS")
        self._fuse_post = _env_flag("FDTD_B200_FUSE")
        # temporally fused E+H steps in run(): 0 never, 1 wherever legal, 2 (default) where also faster (engine.py)
        self._fuse_eh = int(os.environ.get("FDTD_B200_FUSE_EH", "2") or 0)
        self._E2 = self._H2 = None

    # ----------------------------------------------------------------------------- materials
    def _material(self, value, what):
        """-> (background inverse, 3 python floats in the grid dtype; SoA array or None).
        Accepts what the reference accepts (fdtd/grid.py:141-151): a scalar, or an array of shape
        (Nx,Ny,Nz), (Nx,Ny,Nz,1), (Nx,Ny,Nz,3) or anything broadcastable against (Nx,Ny,Nz,3)."""
        if bd.is_complex(value):
            raise NotImplementedError(f"complex {what} is not supported by the CUDA engine")
        if bd.is_array(value) and len(value.shape) == 3:
            value = value[:, :, :, None]
        if torch.is_tensor(value):
            v = value.detach().to(dtype=self._dtype)
        else:
            # via numpy: a python float must enter as float64 (torch.as_tensor would make it float32)
            v = torch.from_numpy(np.asarray(value, dtype=np.float64)).to(dtype=self._dtype)
        ones = torch.ones(3, dtype=self._dtype, device=v.device)
        if v.numel() == 1 or (v.dim() >= 1 and v.numel() == 3 and v.shape[-1] == 3):
            inv = (ones / v.reshape(-1)).cpu()
            return [float(x) for x in inv.double()], None
        full = torch.broadcast_to(v, (self.Nx, self.Ny, self.Nz, 3))
        local = full[self._part.x0:self._part.x1].to(bd.device)
        arr = (torch.ones_like(local) / local).permute(3, 0, 1, 2).contiguous()
        bg = [float(x) for x in arr[:, 0, 0, 0].double().cpu()]
        return bg, arr

    def _materialize(self, which):
        """create the full inverse-material array of a so-far homogeneous grid."""
        name, bgname = ("_inv_eps", "_bg_inv_eps") if which == "eps" else ("_inv_mu", "_bg_inv_mu")
        if getattr(self, name) is None:
            bg = torch.tensor(getattr(self, bgname), dtype=torch.float64).to(self._dtype).to(bd.device)
            arr = bg.view(3, 1, 1, 1).expand(3, self._part.nx, self.Ny, self.Nz).contiguous()
            setattr(self, name, arr)
            self._registration_count += 1
        return getattr(self, name)

    @property
    def inverse_permittivity(self):
        """(Nx,Ny,Nz,3) view (local slab when sharded); zero inside objects like the reference's."""
        return self._materialize("eps").permute(1, 2, 3, 0)

    @inverse_permittivity.setter
    def inverse_permittivity(self, value):
        self._materialize("eps").permute(1, 2, 3, 0).copy_(torch.as_tensor(value, device=bd.device))

    @property
    def inverse_permeability(self):
        return self._materialize("mu").permute(1, 2, 3, 0)

    @inverse_permeability.setter
    def inverse_permeability(self, value):
        self._materialize("mu").permute(1, 2, 3, 0).copy_(torch.as_tensor(value, device=bd.device))

    # -------------------------------------------------------------------------------- fields
    def _local_view(self, F):
        return F[:, 1:-1].permute(1, 2, 3, 0)

    def _get_field(self, F):
        if self._engine is not None:
            self._engine.quiesce()
        if not self._part.sharded:
            return self._local_view(F)
        return all_gather_slabs(self._part, self._local_view(F).contiguous(), 0)

    def _set_field(self, F, value):
        view = self._local_view(F)
        if torch.is_tensor(value) and value.data_ptr() == view.data_ptr() and value.shape == view.shape:
            return                                  # `grid.E *= 0` re-assigns the view it was given
        value = torch.as_tensor(value, device=bd.device) if not torch.is_tensor(value) else value.to(bd.device)
        if self._part.sharded and value.dim() == 4 and value.shape[0] == self.Nx:
            value = value[self._part.x0:self._part.x1]
        view.copy_(value)
        if self._engine is not None and self._engine._halo is not None:
            if self._engine._p2p:
                self._engine._p2p_refresh()
            else:
                self._engine._halo.refresh()

    @property
    def E(self):
        return self._get_field(self._E)

    @E.setter
    def E(self, value):
        self._set_field(self._E, value)

    @property
    def H(self):
        return self._get_field(self._H)

    @H.setter
    def H(self, value):
        self._set_field(self._H, value)

    @property
    def E_local(self):
        return self._local_view(self._E)

    @property
    def H_local(self):
        return self._local_view(self._H)

    # ------------------------------------------------------------------- index handling (G1)
    def _handle_distance(self, distance) -> int:
        """distance -> cells; anything that is not a python int is metres (fdtd/grid.py:171-175)."""
        if not isinstance(distance, int):
            return int(float(distance) / self.grid_spacing + 0.5)
        return distance

    def _handle_time(self, time) -> int:
        if not isinstance(time, int):
            return int(float(time) / self.time_step + 0.5)
        return time

    def _handle_tuple(self, shape):
        if len(shape) != 3:
            raise ValueError(
                f"invalid grid shape {shape}\ngrid shape should be a 3D tuple containing floats or ints")
        x, y, z = shape
        return self._handle_distance(x), self._handle_distance(y), self._handle_distance(z)

    def _handle_slice(self, s: slice) -> slice:
        conv = lambda v: self._handle_distance(v) if isinstance(v, float) else v
        return slice(conv(s.start), conv(s.stop), conv(s.step))

    def _handle_single_key(self, key):
        try:
            len(key)
            return [self._handle_distance(k) for k in key]
        except TypeError:
            if isinstance(key, slice):
                return self._handle_slice(key)
            return [self._handle_distance(key)]

    @property
    def x(self):
        return self.Nx * self.grid_spacing

    @property
    def y(self):
        return self.Ny * self.grid_spacing

    @property
    def z(self):
        return self.Nz * self.grid_spacing

    @property
    def shape(self):
        return (self.Nx, self.Ny, self.Nz)

    @property
    def time_passed(self) -> float:
        return self.time_steps_passed * self.time_step

    # ------------------------------------------------------------------------------ stepping
    def _ready(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self)
        elif self._engine.stale():
            self._engine.flush_detectors()
            self._engine.bake()
        return self._engine

    def run(self, total_time, progress_bar: bool = True):
        """run the simulation for `total_time` (int: steps, float: seconds), fdtd/grid.py:250-265."""
        if isinstance(total_time, float):
            total_time /= self.time_step
        n = max(0, int(total_time))
        eng = self._ready()
        bar = None
        if progress_bar:
            from tqdm import tqdm
            bar = tqdm(total=n)
        try:
            eng.run(self.time_steps_passed, n, bar)       # advances time_steps_passed chunk by chunk
        finally:
            if bar is not None:
                bar.close()

    def step(self):
        """one full step: update_E, update_H, count (fdtd/grid.py:267-273)."""
        self.update_E()
        self.update_H()
        self.time_steps_passed += 1

    def update_E(self):
        """fused E half-step (fdtd/grid.py:275-299)."""
        self._ready().update_E(self.time_steps_passed)

    def update_H(self):
        """fused H half-step (fdtd/grid.py:301-325)."""
        self._ready().update_H(self.time_steps_passed)

    def reset(self):
        """fields and the step counter to zero; PML state and detector histories are kept, as in
        the reference (fdtd/grid.py:327-331)."""
        self._H.mul_(0.0)
        self._E.mul_(0.0)
        self.time_steps_passed *= 0

    def synchronize(self):
        if self._E.is_cuda:
            torch.cuda.synchronize(self._E.device)

    # --------------------------------------------------------------------------- registration
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
        attr._register_grid(
            grid=self,
            x=self._handle_single_key(x),
            y=self._handle_single_key(y),
            z=self._handle_single_key(z),
        )
        self._registration_count += 1

    def _register_name(self, thing):
        if thing.name is not None:
            if not hasattr(self, thing.name):
                setattr(self, thing.name, thing)
            else:
                raise ValueError(f"The grid already has an attribute with name {thing.name}")

    # ------------------------------------------------ simulation I/O (fdtd/grid.py:407-439, 490-514)
    def save_simulation(self, sim_name=None):
        """create `fdtd_output/fdtd_output_<timestamp>[ (sim_name)]` and remember it (fdtd/grid.py:407-439)."""
        from datetime import datetime
        os.makedirs("fdtd_output", exist_ok=True)
        now = datetime.now()
        full_sim_name = f"{now.year}-{now.month}-{now.day}-{now.hour}-{now.minute}-{now.second}"
        if sim_name is not None:
            full_sim_name = full_sim_name + " (" + sim_name + ")"
        self.folder = os.path.abspath(os.path.join("fdtd_output", "fdtd_output_" + full_sim_name))
        self.full_sim_name = full_sim_name
        os.makedirs(self.folder, exist_ok=True)
        return self.folder

    def save_data(self):
        """detector readings -> `<folder>/detector_readings.npz`, keys "<name> (E)" / "<name> (H)" per detector
        as in the reference (fdtd/grid.py:490-514); the device rings are drained first."""
        import numpy as np
        if self.folder is None:
            raise Exception("Save location not initialized. Please read about 'fdtd.Grid.saveSimulation()' or "
                            "try running 'grid.saveSimulation()'.")
        dic = {}
        for detector in self.detectors:
            for key, values in detector.detector_values().items():
                dic[f"{detector.name} ({key})"] = np.asarray(values)
        np.savez(os.path.join(self.folder, "detector_readings"), **dic)

    def generate_video(self, delete_frames=False):
        """frames written by `visualize(save=True, folder=grid.folder, index=n)` -> one mp4 in the simulation
        folder, through the `ffmpeg` executable (fdtd/grid.py:441-488).  Returns the file name."""
        import glob
        import subprocess
        if self.folder is None:
            raise Exception("Save location not initialized. Please read about 'fdtd.Grid.saveSimulation()' or "
                            "try running 'grid.saveSimulation()'.")
        name = "fdtd_sim_video_" + self.full_sim_name + ".mp4"
        cmd = ["ffmpeg", "-y", "-framerate", "8", "-i", "file%04d.png", "-r", "30", "-pix_fmt", "yuv420p", name]
        try:
            subprocess.check_call(cmd, cwd=self.folder)
        except (FileNotFoundError, subprocess.CalledProcessError) as exc:
            raise RuntimeError("Error when calling ffmpeg. Is ffmpeg installed and available in your path?") from exc
        if delete_frames:
            for frame in glob.glob(os.path.join(self.folder, "file*.png")):
                os.remove(frame)
        return name

    def promote_dtypes_to_complex(self):
        raise NotImplementedError("complex fields are not supported by the CUDA engine")

    def __repr__(self):
        return (f"{self.__class__.__name__}(shape=({self.Nx},{self.Ny},{self.Nz}), "
                f"grid_spacing={self.grid_spacing:.2e}, courant_number={self.courant_number:.2f})")

    def __str__(self):
        s = repr(self) + "\n"
        for title, items in (("sources", self.sources), ("detectors", self.detectors),
                             ("boundaries", self.boundaries), ("objects", self.objects)):
            if items:
                s = s + f"\n{title}:\n"
                for item in items:
                    s += str(item)
        return s
