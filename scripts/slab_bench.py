"""Per-step time of an x-sharded c4-structured grid with thin slabs (development tool; run under torchrun).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/slab_bench.py 256 [steps]

The grid is (NX, 1024, 1024) float32 with six PMLs: NX / world planes per rank reproduce what one rank of eight holds
of the 1024^3 workload on a cheaper box.  Prints ms per step (CUDA events, max over ranks) of grid.run(steps).
Knobs come from the environment: FDTD_B200_FUSE_EH (0: two half-steps), FDTD_B200_FUSE_SPLIT (0: one launch)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import fdtd_b200 as fd
    from bench import build_c4
    fd.set_backend("cuda.float32")
    g = build_c4(fd, (nx, 1024, 1024), balance=True)
    if "FDTD_B200_FUSE_EH" in os.environ:
        g._fuse_eh = int(os.environ["FDTD_B200_FUSE_EH"])
    g.run(6, progress_bar=False)
    g._engine.flush_detectors()
    best = None
    for _ in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.run(steps, progress_bar=False)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / steps], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = float(ms.item()) if best is None else min(best, float(ms.item()))
        g._engine.flush_detectors()
    if dist.get_rank() == 0:
        import ctypes
        eng = g._engine
        fused = bool(eng._p2p) and eng.lib.fdtd_fuse_eh_sharded_active(ctypes.byref(eng.desc), ctypes.byref(eng._p2p.h)) == 1
        print(f"{nx}x1024x1024 on {dist.get_world_size()} GPUs ({nx // dist.get_world_size()} planes each), fused steps "
              f"{fused}, FUSE_SPLIT={os.environ.get('FDTD_B200_FUSE_SPLIT', '1')}: {best:.4f} ms per step "
              f"({nx * 1024 * 1024 / best / 1e6:.1f} Gcell/s)", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
