"""BASELINE.json configs at full size through the public API on one B200 (SURVEY.md section 8d inputs).

    python scripts/bench_configs.py [c1 c2 c3 c3b c4 c5slab]

Prints one JSON line per config: Mcell-updates/s, ms/step, algorithmic bytes per cell-step and the
fraction of the measured HBM peak.  Timing: CUDA events around grid.run(steps) after a warm-up run
(detector flush and waveform upload inside the region, like bench.py's e2e).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdtd_b200 as fd  # noqa: E402
from bench import measured_peak  # noqa: E402

SPACING = 77.5e-9


def six_pml(g, t=10):
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, -t:, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[:, :, -t:] = fd.PML()


def c1():
    """configs[0]: README quick-start 161x97x1, float64, 1000 steps."""
    fd.set_backend("cuda.float64")
    g = fd.Grid(shape=(25e-6, 15e-6, 1), grid_spacing=155e-9)
    g[11:32, 30:84, 0] = fd.Object(permittivity=1.7 ** 2, name="object")
    g[7.5e-6:8.0e-6, 11.8e-6:13.0e-6, 0] = fd.LineSource(period=1550e-9 / (3e8), name="source")
    g[12e-6, :, 0] = fd.LineDetector(name="detector")
    g[0:10, :, :] = fd.PML()
    g[-10:, :, :] = fd.PML()
    g[:, 0:10, :] = fd.PML()
    g[:, -10:, :] = fd.PML()
    g[:, :, 0] = fd.PeriodicBoundary(name="zbounds")
    return g, 1000, 8, None


def c2():
    """configs[1]: 256^3 float64, six PMLs, PointSource + BlockDetector, 2000 steps."""
    fd.set_backend("cuda.float64")
    n = 256
    g = fd.Grid(shape=(n, n, n), grid_spacing=SPACING)
    six_pml(g)
    g[128, 128, 128] = fd.PointSource(period=20)
    g[138:140, 128:130, 128:130] = fd.BlockDetector()
    return g, 2000, 8, 18 + 8 * 60 / n


def _lens(n0=260, n1=324, a=128, b=384):
    i = np.arange(n0, n1)[:, None, None]
    j = np.arange(a, b)[None, :, None]
    k = np.arange(a, b)[None, None, :]
    mask = ((j - 256) ** 2 + (k - 256) ** 2 + (i - 164) ** 2) <= 160 ** 2
    P = np.ones((n1 - n0, b - a, b - a, 3))
    P[mask] = (2.25, 2.25, 2.5)
    return P


def c3(as_grid_permittivity=False):
    """configs[2]: 512^3 float32, AbsorbingObject + anisotropic lens, six PMLs, PlaneSource, 500 steps."""
    fd.set_backend("cuda.float32")
    n = 512
    if as_grid_permittivity:
        eps = np.ones((n, n, n, 3), dtype=np.float32)
        eps[260:324, 128:384, 128:384] = _lens()
        g = fd.Grid(shape=(n, n, n), grid_spacing=SPACING, permittivity=eps)
    else:
        g = fd.Grid(shape=(n, n, n), grid_spacing=SPACING)
    six_pml(g)
    g[60, :, :] = fd.PlaneSource(period=20, polarization="z")
    g[150:200, 100:412, 100:412] = fd.AbsorbingObject(permittivity=2.5, conductivity=1.5e4)
    if not as_grid_permittivity:
        g[260:324, 128:384, 128:384] = fd.AnisotropicObject(permittivity=_lens())
    g[20:492, 256, 256] = fd.LineDetector()
    # words/cell-step: 18 + PML 8*60/512 + absorber (3.6 % of cells) 6 words + lens (3.1 %) 3 words
    absorber = 50 * 312 * 312 / n ** 3
    lens = 64 * 256 * 256 / n ** 3
    words = 18 + 8 * 60 / n + 6 * absorber + 3 * lens     # (c3b: only tiles that differ from the background stream eps^-1)
    return g, 500, 4, words


def c4():
    """configs[3]: 1024^3 float32, six PMLs, PointSource, LineDetector, 200 steps (the bench workload)."""
    fd.set_backend("cuda.float32")
    n = 1024
    g = fd.Grid(shape=(n, n, n), grid_spacing=SPACING)
    six_pml(g)
    g[512, 512, 512] = fd.PointSource(period=20)
    g[516, 512, 12:1012] = fd.LineDetector()
    return g, 200, 4, 18 + 8 * 60 / n


def c5slab():
    """configs[4], one GPU's share (256 x 1024 x 1024 of the 2048-long waveguide): periodic y/z, GRIN
    object over the slab, PlaneSource; x-PML on the low side only (the rank-0 slab)."""
    fd.set_backend("cuda.float32")
    nx, n = 256, 1024
    g = fd.Grid(shape=(nx, n, n), grid_spacing=SPACING)
    g[0:10, :, :] = fd.PML()
    g[:, 0, :] = fd.PeriodicBoundary()
    g[:, :, 0] = fd.PeriodicBoundary()
    ramp = (1.0 + 1.25 * np.arange(n) / (n - 1.0)).reshape(1, n, 1)
    g[128:256, :, :] = fd.Object(permittivity=ramp)
    g[100, :, :] = fd.PlaneSource(period=20, polarization="z")
    g[20:250, 512, 512] = fd.LineDetector()
    words = 18 + 8 * 10 / nx + 3 * 0.5
    return g, 100, 4, words


def c5(nx=2048):
    """configs[4]: 2048 x 1024 x 1024 float32 waveguide, PML on x, periodic y and z, GRIN medium over the middle
    half, PlaneSource; x-sharded over the ranks of the job (torchrun).  `c5weak` gives every rank 256 planes."""
    fd.set_backend("cuda.float32")
    n = 1024
    g = fd.Grid(shape=(nx, n, n), grid_spacing=SPACING)
    g[0:10, :, :] = fd.PML()
    g[-10:, :, :] = fd.PML()
    g[:, 0, :] = fd.PeriodicBoundary()
    g[:, :, 0] = fd.PeriodicBoundary()
    ramp = (1.0 + 1.25 * np.arange(n) / (n - 1.0)).reshape(1, n, 1)
    g[nx // 4:3 * nx // 4, :, :] = fd.Object(permittivity=ramp)
    g[100, :, :] = fd.PlaneSource(period=20, polarization="z")
    g[20:nx - 20, 512, 512] = fd.LineDetector()
    words = 18 + 8 * 20 / nx + 3 * 0.5
    return g, 100, 4, words


def c5weak():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return c5(256 * world)


CONFIGS = {"c1": c1, "c2": c2, "c3": c3, "c3b": lambda: c3(True), "c4": c4, "c5slab": c5slab, "c5": c5,
           "c5weak": c5weak}


def main():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    names = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5slab"]
    peak, src = measured_peak()
    for name in names:
        t0 = time.perf_counter()
        g, steps, warm, words = CONFIGS[name]()
        g._x_chunk = int(os.environ.get("X_CHUNK", "0"))
        if os.environ.get("TUNE_LIB"):          # development: time a build variant from scripts/tune.py
            from fdtd_b200 import _capi
            fd.backend.lib = _capi.bind(os.environ["TUNE_LIB"])
        g.run(warm, progress_bar=False)
        for det in g.detectors:
            _ = det.E
        torch.cuda.synchronize()
        setup_s = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.run(steps, progress_bar=False)
        traces = [(det.E, det.H) for det in g.detectors]
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        cells = g.Nx * g.Ny * g.Nz
        w = 4 if g._dtype is torch.float32 else 8
        rec = {"config": name, "grid": [g.Nx, g.Ny, g.Nz], "dtype": str(g._dtype).split(".")[-1], "steps": steps,
               "ms_per_step": ms / steps, "Mcell_per_s": cells * steps / (ms * 1e-3) / 1e6,
               "steps_per_s": steps / (ms * 1e-3), "setup_s": round(setup_s, 2),
               "fused_post": bool(g._engine.lib.fdtd_post_is_fused(g._engine.desc)),
               "graphs": bool(g._engine.desc.use_graphs), "detector_samples": len(traces[0][0]) if traces else 0,
               "E_absmax": float(g.E_local.abs().max()), "n_gpus": world,
               "halo": ("p2p" if g._engine._p2p else "nccl") if world > 1 else None}
        if words:
            rec["bytes_per_cell_step"] = w * words
            rec["hbm_frac_of_measured"] = (w * words * cells * steps / (ms * 1e-3) / 1e9) / (peak * world)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del g, traces
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
