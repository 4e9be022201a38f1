"""NVLink traffic of the half-step kernel that also pushes its boundary plane (development tool, ONE process, two
GPUs): the slab lives on cuda:0, the "neighbour's ghost planes" on cuda:1 with peer access enabled, so that ncu on this
single process sees the peer stores of `halfstep_kernel<..., HAS_PUSH = 1, ...>` as NVLink bytes.

    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:halfstep_kernel python scripts/nvlink_push_probe.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fdtd_b200 as fd  # noqa: E402
from bench import build_c4  # noqa: E402
from fdtd_b200 import _capi  # noqa: E402


def main():
    assert torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1)
    torch.cuda.set_device(0)
    fd.set_backend("cuda.float32")
    n = (128, 1024, 1024)                      # what one rank of eight holds of the 1024^3 workload
    g = build_c4(fd, n)
    g._fuse_eh = 0
    g.run(2, progress_bar=False)
    eng, lib = g._engine, fd.backend.lib
    ghost = torch.zeros((2, n[1], n[2]), dtype=torch.float32, device="cuda:1")
    probe = torch.ones(8, device="cuda:0")
    ghost[0, 0, :8].copy_(probe)               # a peer copy: makes torch enable peer access 0 -> 1
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = eng.desc
    for field, plane in ((0, 0), (1, d.Nx - 1)):
        for _ in range(3):
            _capi.check(lib, lib.fdtd_halfstep_push(C.byref(d), field, plane, plane + 1, g.time_steps_passed, 0,
                                                    C.c_void_p(ghost[0].data_ptr()), C.c_void_p(ghost[1].data_ptr()), st))
    torch.cuda.synchronize(0)
    F = g._E if False else g._H
    want_y, want_z = g._H[1, d.Nx], g._H[2, d.Nx]          # storage index = local plane + 1
    got = ghost.to("cuda:0")
    print("pushed plane equals the slab's boundary plane:", bool(torch.equal(got[0], want_y) and torch.equal(got[1], want_z)),
          "| bytes per push:", 2 * n[1] * n[2] * 4)


if __name__ == "__main__":
    main()
