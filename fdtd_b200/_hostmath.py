"""Host-side coefficient arithmetic (registration time only, never per step).

The few transcendental / reduction results that feed the kernels (PML b and c profiles, the
LineSource gaussian profile) are computed on the host with the SAME library the matching
reference backend uses, so that the device tables are bit-identical to the reference's:
numpy for float64 (reference default backend), torch-CPU for float32 (the only way the
reference computes in true float32).  This is a table builder -- a few
dozen numbers -- not a compute path.
"""
import numpy as np
import torch


class HostLib:
    def __init__(self, dtype):
        self.torch_dtype = dtype
        self.use_torch = dtype is torch.float32
        self.np_dtype = np.float32 if dtype is torch.float32 else np.float64

    def zeros(self, n):
        return torch.zeros(n, dtype=torch.float32) if self.use_torch else np.zeros(n, dtype=np.float64)

    def arange(self, a, b, s):
        if self.use_torch:
            return torch.arange(a, b, s, dtype=torch.float32)
        return np.arange(a, b, s, dtype=np.float64)

    def asarray(self, a):
        if self.use_torch:
            return torch.tensor(np.asarray(a), dtype=torch.float32)
        return np.array(a, dtype=np.float64)

    def exp(self, a):
        return torch.exp(a) if self.use_torch else np.exp(a)

    def to_device(self, a, device):
        """host table -> device tensor of the grid dtype (exact: same dtype)."""
        if not torch.is_tensor(a):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(device=device, dtype=self.torch_dtype).contiguous()


def scalar_in_dtype(value, dtype):
    """python float rounded to the grid dtype (what `python_float * array` does in the reference)."""
    return float(torch.tensor(value, dtype=torch.float64).to(dtype).item())
