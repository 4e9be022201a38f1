"""x-slab partition of a Grid across the ranks of a torch.distributed job, and the one-plane
halo exchange the Yee update needs each half-step.

One process per GPU.  The stencil reaches one cell along x:
  * the E half-step at local plane 0 needs (Hy, Hz) of the LEFT neighbour's last plane,
  * the H half-step at the last local plane needs (Ey, Ez) of the RIGHT neighbour's first plane.
Field storage carries one ghost x-plane at each end (include/fdtd_b200.h); an exchange is two
contiguous Ny*Nz planes per neighbour, sent with NCCL send/recv on a side stream so that it
overlaps with the bulk of the following half-step.  The `gloo` backend (CPU tensors) is
supported for the host-logic tests only.
"""
import os

import torch
import torch.distributed as dist


class Partition:
    """contiguous, balanced x-slabs: rank r owns global planes [x0, x1)."""

    def __init__(self, Nx, shard="auto", plane_cost=None):
        active = (shard not in (None, False) and dist.is_available() and dist.is_initialized()
                  and dist.get_world_size() > 1)
        if shard is True and not active:
            raise RuntimeError("shard=True needs an initialised torch.distributed job with world_size > 1")
        self.Nx = Nx
        self.rank = dist.get_rank() if active else 0
        self.world = dist.get_world_size() if active else 1
        if self.world > Nx:
            raise ValueError(f"cannot shard Nx={Nx} planes over {self.world} ranks")
        self._cuts = self._balanced_cuts(plane_cost)
        self.x0, self.x1 = self.bounds(self.rank)

    def _balanced_cuts(self, plane_cost):
        """slab boundaries: equal plane counts, or -- given a relative cost per x-plane (planes inside an
        x-PML move 13 words per cell and half-step instead of 9) -- equal cumulative cost."""
        n, p = self.Nx, self.world
        if plane_cost is None or p == 1:
            base, rem = divmod(n, p)
            return [r * base + min(r, rem) for r in range(p)] + [n]
        cost = [float(c) for c in plane_cost]
        if len(cost) != n or min(cost) <= 0:
            raise ValueError("plane_cost needs one positive entry per x-plane")
        total, cuts, acc, i = sum(cost), [0], 0.0, 0
        for r in range(1, p):
            target = total * r / p
            while i < n and acc + cost[i] / 2 < target:
                acc += cost[i]
                i += 1
            i = max(i, cuts[-1] + 1)                 # at least one plane per rank
            i = min(i, n - (p - r))
            acc = sum(cost[:i])
            cuts.append(i)
        return cuts + [n]

    def bounds(self, r):
        return self._cuts[r], self._cuts[r + 1]

    @property
    def nx(self):
        return self.x1 - self.x0

    @property
    def sharded(self):
        return self.world > 1

    def local_range(self, g0, g1):
        """global half-open x-range -> local half-open range clipped to this slab (may be empty)."""
        a, b = max(g0, self.x0), min(g1, self.x1)
        if a >= b:
            return 0, 0
        return a - self.x0, b - self.x0

    def owner(self, gx):
        for r in range(self.world):
            a, b = self.bounds(r)
            if a <= gx < b:
                return r
        raise IndexError(gx)


class HaloExchange:
    """asynchronous one-plane exchanges of (Ey,Ez) to the left and (Hy,Hz) to the right."""

    def __init__(self, part, E, H, wrap_H=False):
        self.part, self.E, self.H = part, E, H      # storage tensors (3, nx+2, Ny, Nz)
        self.cuda = E.is_cuda
        self.stream = torch.cuda.Stream(device=E.device) if self.cuda else None
        # wrap_H: the last slab's last H plane also goes into the FIRST slab's low ghost plane (a CurrentDetector on
        # global plane x = 0 reads H[-1], fdtd/detectors.py:432-447)
        self.wrap_H = wrap_H

    def _ops(self, F, to_left):
        p, n = self.part, self.part.nx
        ops = []
        for c in (1, 2):                             # only the y and z components cross an x face
            if to_left:
                if p.rank > 0:
                    ops.append(dist.P2POp(dist.isend, F[c, 1], p.rank - 1))
                if p.rank < p.world - 1:
                    ops.append(dist.P2POp(dist.irecv, F[c, n + 1], p.rank + 1))
            else:
                if p.rank < p.world - 1:
                    ops.append(dist.P2POp(dist.isend, F[c, n], p.rank + 1))
                if p.rank > 0:
                    ops.append(dist.P2POp(dist.irecv, F[c, 0], p.rank - 1))
                if self.wrap_H and F is self.H:
                    if p.rank == p.world - 1:
                        ops.append(dist.P2POp(dist.isend, F[c, n], 0))
                    if p.rank == 0:
                        ops.append(dist.P2POp(dist.irecv, F[c, 0], p.world - 1))
        return ops

    def start(self, field):
        """begin the exchange that follows a half-step of `field` ('E' -> left, 'H' -> right).
        Returns a handle for wait()."""
        ops = self._ops(self.E if field == "E" else self.H, to_left=(field == "E"))
        if not ops or os.environ.get("FDTD_B200_SKIP_HALO"):     # (timing experiments only: wrong results)
            return None
        if self.cuda:
            main = torch.cuda.current_stream(self.E.device)
            self.stream.wait_stream(main)
            with torch.cuda.stream(self.stream):
                for r in dist.batch_isend_irecv(ops):
                    r.wait()                         # NCCL: orders the side stream, does not block the host
                ev = torch.cuda.Event()
                ev.record(self.stream)
            return ev                                # (the boundary plane of the next half-step runs on
                                                     #  this same stream, right behind the exchange)
        return [op.op(op.tensor, op.peer) for op in ops]   # gloo: plain isend / irecv requests

    def wait(self, handle):
        if handle is None:
            return
        if self.cuda:
            torch.cuda.current_stream(self.E.device).wait_event(handle)
        else:
            for r in handle:
                r.wait()

    def refresh(self):
        """synchronous exchange of both fields (after the user wrote E or H)."""
        self.wait(self.start("E"))
        self.wait(self.start("H"))


class WrapExchange:
    """The plane copy of a periodic x boundary on an x-sharded grid (fdtd/boundaries.py:184-195): after the E
    update E[0] = E[-1] moves the last plane of the last slab into the first plane of the first slab, after the
    H update H[-1] = H[0] goes the other way.  All three components travel, through one staging buffer, as a
    send / recv pair of the process group (NCCL or gloo) ordered on the caller's stream; the other ranks do
    nothing."""

    def __init__(self, part, E, H):
        self.part, self.E, self.H = part, E, H      # storage tensors (3, nx+2, Ny, Nz)
        self.first, self.last = part.rank == 0, part.rank == part.world - 1
        self.stage = E.new_empty((3,) + tuple(E.shape[2:])) if (self.first or self.last) else None

    def run(self, field):
        if not (self.first or self.last):
            return
        F = self.E if field == "E" else self.H
        n = self.part.nx
        sender = self.last if field == "E" else self.first
        src_plane, dst_plane = (n, 1) if field == "E" else (1, n)        # storage index = local plane + 1
        peer = 0 if self.last else self.part.world - 1
        if sender:
            self.stage.copy_(F[:, src_plane])
            dist.send(self.stage, dst=peer)
        else:
            dist.recv(self.stage, src=peer)
            F[:, dst_plane].copy_(self.stage)


class P2PHalo:
    """Halo exchange by direct stores into the neighbour slab's ghost planes over NVLink.

    One process per GPU: every rank exports its E / H storage and a two-word flag array with CUDA IPC
    (include/fdtd_b200.h, fdtd_ipc_export / fdtd_ipc_import) and maps its neighbours'.  The result is one
    `fdtd_halo` struct (`self.h`): peer pointers, flag addresses, running push counts.  The library does the rest
    (fdtd_sharded_halfstep / fdtd_run_sharded): the half-step kernel that computes a slab's boundary plane stores it
    into the neighbour's ghost plane as well (compute + transfer in one kernel); a one-thread kernel then publishes
    the new half-step count in the neighbour's flag (release, system scope) and the neighbour's stream spins on its
    local flag (acquire) before the kernel that consumes the ghost.  No NCCL call, no host synchronisation between
    the processes; counts are monotonic, so ranks may run ahead of each other by at most one half-step (the
    dependency chain of the flags is the back-pressure).  A wait that exceeds FDTD_B200_HALO_TIMEOUT_S (default
    120 s) raises the error word and traps, so a dead neighbour can never turn into a silently wrong result."""

    def __init__(self, part, E, H, lib, E2=None, H2=None):
        import ctypes as C
        from . import _capi
        self.part, self.E, self.H, self.lib = part, E, H, lib
        self.fused_buffers = E2 is not None and H2 is not None     # second field buffers of temporally fused steps
        self.cuda = True
        dev = E.device
        # high priority: the boundary-plane launches and flags are dispatched ahead of the bulk's pending blocks, so the
        # neighbour gets its ghost plane while the bulk is still running
        self.stream = torch.cuda.Stream(device=dev, priority=-1)
        self.flags = torch.zeros(2, dtype=torch.int64, device=dev)     # [0]: E pushes received, [1]: H pushes
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)

        def export(t):
            h = (C.c_char * 64)()
            off = C.c_int64()
            _capi.check(lib, lib.fdtd_ipc_export(C.c_void_p(t.data_ptr()), C.cast(h, C.c_void_p), C.byref(off)))
            return bytes(h.raw), off.value

        mapped = {}        # IPC handle -> mapped base: tensors of one allocation are mapped once

        def open_(rec):
            if rec[0] not in mapped:
                h = (C.c_char * 64).from_buffer_copy(rec[0])
                ptr = C.c_void_p()
                _capi.check(lib, lib.fdtd_ipc_import(C.cast(h, C.c_void_p), 0, C.byref(ptr)))
                mapped[rec[0]] = ptr.value
            return mapped[rec[0]] + rec[1]

        # every step below is collective: a rank that cannot export or import must not leave the others
        # waiting, and either ALL ranks use peer-to-peer ghosts or none does
        try:
            mine = {"E": export(E), "H": export(H), "flags": export(self.flags), "nx": part.nx}
            if self.fused_buffers:
                mine.update(E2=export(E2), H2=export(H2))
        except Exception as exc:
            mine = {"error": str(exc)}
        everyone = [None] * part.world
        dist.all_gather_object(everyone, mine)
        failure = next((r["error"] for r in everyone if "error" in r), None)
        w = E.element_size()
        plane = E.shape[2] * E.shape[3]
        h = self.h = _capi.Halo()
        h.has_left, h.has_right = int(part.rank > 0), int(part.rank < part.world - 1)
        h.flags, h.error = self.flags.data_ptr(), self.err.data_ptr()
        h.side_stream = self.stream.cuda_stream
        h.timeout_ns = int(float(os.environ.get("FDTD_B200_HALO_TIMEOUT_S", "120")) * 1e9)
        if failure is None:
            try:
                if h.has_left:                                      # E plane 0 -> left neighbour's high ghost
                    rec = everyone[part.rank - 1]
                    base, nxl = open_(rec["E"]), rec["nx"]
                    h.left_ghost_y = base + ((1 * (nxl + 2) + nxl + 1) * plane) * w
                    h.left_ghost_z = base + ((2 * (nxl + 2) + nxl + 1) * plane) * w
                    h.left_flag = open_(rec["flags"])
                    if self.fused_buffers and "E2" in rec:
                        base = open_(rec["E2"])
                        h.left_ghost_y2 = base + ((1 * (nxl + 2) + nxl + 1) * plane) * w
                        h.left_ghost_z2 = base + ((2 * (nxl + 2) + nxl + 1) * plane) * w
                if h.has_right:                                     # H last plane -> right neighbour's low ghost
                    rec = everyone[part.rank + 1]
                    base, nxl = open_(rec["H"]), rec["nx"]
                    h.right_ghost_y = base + (1 * (nxl + 2) * plane) * w
                    h.right_ghost_z = base + (2 * (nxl + 2) * plane) * w
                    h.right_flag = open_(rec["flags"]) + 8
                    if self.fused_buffers and "H2" in rec:
                        base = open_(rec["H2"])
                        h.right_ghost_y2 = base + (1 * (nxl + 2) * plane) * w
                        h.right_ghost_z2 = base + (2 * (nxl + 2) * plane) * w
            except Exception as exc:
                failure = str(exc)
        ok = torch.tensor([0 if failure else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            raise RuntimeError(failure or "a neighbour rank could not map this rank's memory")

    def check(self):
        """raise if a halo wait timed out (call where the host synchronises anyway: ring flushes)"""
        if int(self.err.item()) != 0:
            raise RuntimeError("fdtd_b200: peer-to-peer halo wait timed out (a neighbour rank stopped stepping)")


def all_gather_slabs(part, local, dim):
    """concatenate per-rank slabs of differing thickness along `dim` (convenience accessor)."""
    sizes = [part.bounds(r)[1] - part.bounds(r)[0] for r in range(part.world)]
    nmax = max(sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = nmax
    buf = local.new_zeros(pad_shape)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    out = [torch.empty_like(buf) for _ in range(part.world)]
    dist.all_gather(out, buf)
    return torch.cat([o.narrow(dim, 0, s) for o, s in zip(out, sizes)], dim=dim)
