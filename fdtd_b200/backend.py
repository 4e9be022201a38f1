"""Backend selection, mirroring the reference's `fdtd.backend` / `fdtd.set_backend`
(fdtd/backend.py:363-439) with ONE engine: hand-written sm_100a CUDA kernels behind the C ABI
of include/fdtd_b200.h.  Accepted names:

    "cuda"            float64 storage and arithmetic (the reference's default precision)
    "cuda.float64"
    "cuda.float32"    true float32 storage and arithmetic
    "cuda.float32x"   float32 storage of the state (E, H, CPML psi, detector rings) with float64 arithmetic and
                      coefficients: every value is widened where it is used and rounded once where it is stored.
                      Same memory footprint and HBM traffic as "cuda.float32"; stays within 1e-5 of the reference
                      (which computes in float64 whatever its backend name says) over thousands of steps
    "torch.cuda[.float32|.float64]"   accepted as aliases of the above for drop-in scripts

Every other reference name ("numpy", "torch", ...) raises: this package has no CPU path.
Unlike the reference -- whose ".float32" names still compute in float64
(fdtd/backend.py:43-49,79-89,281-284) -- "cuda.float32" is float32.

The `backend` singleton also carries the small array-function surface the reference's
user-facing code uses (`bd.array`, `bd.zeros`, `bd.numpy`, ...), on CUDA tensors.
"""
import numpy as _np
import torch

from . import _capi


class Backend:
    """array helpers on the engine's device + the loaded C-ABI library."""

    pi = _np.pi
    int = torch.int64

    def __init__(self):
        self.name = None
        self.device = None
        self.float = torch.float64        # arithmetic / coefficient dtype (what `bd.zeros`, `bd.array` ... create)
        self.storage = torch.float64      # dtype of the state: fields, CPML psi, detector rings
        self.lib = None

    # --- state ------------------------------------------------------------------------------
    @property
    def ready(self):
        return self.lib is not None

    def require(self):
        if not self.ready:
            set_backend("cuda")
        return self

    @property
    def complex(self):
        return torch.complex64 if self.float is torch.float32 else torch.complex128

    def __repr__(self):
        return f"CudaBackend({self.name})"

    # --- array surface (fdtd/backend.py:93-355) -----------------------------------------------
    def zeros(self, shape, dtype=None, **kw):
        return torch.zeros(shape, dtype=dtype or self.float, device=self.device, **kw)

    def ones(self, shape, dtype=None, **kw):
        return torch.ones(shape, dtype=dtype or self.float, device=self.device, **kw)

    def zeros_like(self, a):
        return torch.zeros_like(a)

    def array(self, arr, dtype=None):
        dtype = dtype or self.float
        if torch.is_tensor(arr):
            return arr.clone().to(device=self.device, dtype=dtype)
        return torch.tensor(_np.asarray(arr), device=self.device, dtype=dtype)

    def asarray(self, arr):
        return torch.as_tensor(arr, device=self.device)

    def arange(self, *a, **kw):
        return torch.arange(*a, device=self.device, **kw)

    def linspace(self, start, stop, num=50, endpoint=True):
        return torch.as_tensor(_np.linspace(start, stop, num, endpoint=endpoint), device=self.device)

    def numpy(self, arr):
        if torch.is_tensor(arr):
            return arr.detach().cpu().numpy()
        return _np.asarray(arr)

    @staticmethod
    def is_array(arr):
        return isinstance(arr, _np.ndarray) or torch.is_tensor(arr)

    @staticmethod
    def is_complex(x):
        if isinstance(x, complex):
            return True
        if torch.is_tensor(x):
            return torch.is_complex(x)
        return isinstance(x, _np.ndarray) and _np.iscomplexobj(x)

    exp = staticmethod(torch.exp)
    sin = staticmethod(torch.sin)
    cos = staticmethod(torch.cos)
    sum = staticmethod(torch.sum)
    max = staticmethod(torch.max)
    stack = staticmethod(torch.stack)
    squeeze = staticmethod(torch.squeeze)
    reshape = staticmethod(torch.reshape)
    bmm = staticmethod(torch.bmm)
    broadcast_arrays = staticmethod(torch.broadcast_tensors)
    broadcast_to = staticmethod(torch.broadcast_to)
    divide = staticmethod(torch.div)
    fftfreq = staticmethod(_np.fft.fftfreq)
    fft = staticmethod(torch.fft.fft)
    pad = staticmethod(torch.nn.functional.pad)

    @staticmethod
    def transpose(arr, axes=None):
        if axes is None:
            axes = tuple(range(arr.dim() - 1, -1, -1))
        return arr.permute(*axes)


backend = Backend()

# name -> (arithmetic / coefficient dtype, storage dtype of the state)
_NAMES = {
    "cuda": (torch.float64, torch.float64), "cuda.float64": (torch.float64, torch.float64),
    "cuda.float32": (torch.float32, torch.float32), "cuda.float32x": (torch.float64, torch.float32),
    "torch.cuda": (torch.float64, torch.float64), "torch.cuda.float64": (torch.float64, torch.float64),
    "torch.cuda.float32": (torch.float32, torch.float32),
}


def set_backend(name: str):
    """Select the CUDA engine and its precision (before building a Grid, as in the reference)."""
    if name not in _NAMES:
        known = ("numpy", "torch")
        if any(name == k or name.startswith(k + ".") for k in known):
            raise ValueError(
                f"backend '{name}' is a CPU / PyTorch-eager backend of the reference; fdtd_b200 only "
                f"provides the CUDA engine: {sorted(_NAMES)} (no CPU fallback)")
        raise ValueError(f"Unknown backend '{name}'. Available backends: {sorted(_NAMES)}")
    if not torch.cuda.is_available():
        raise RuntimeError("fdtd_b200: no CUDA device available.\nThe engine has no CPU fallback.")
    lib = _capi.load()  # RuntimeError if the extension is not built
    backend.lib = lib
    backend.device = torch.device("cuda", torch.cuda.current_device())
    backend.float, backend.storage = _NAMES[name]
    backend.name = name
    return backend
