#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of the fused 3-D Yee step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size S] [--dtype float32]

Workload (config.workload): BASELINE.json configs[3] -- 3-D S^3 (default 1024^3) float32 grid, 10-cell
PML on all six faces, PointSource(period=20) at the centre, one LineDetector; strong scaling over x-slabs
for N > 1 (one process per GPU, launched by torchrun; peer-to-peer ghost-plane stores over NVLink, NCCL fallback).  A "step" is one full
E+H update of the whole grid including PML, source and detector work.

Printed JSON (one line, rank 0):
  value        Mcell-updates/s over all GPUs, fields resident in HBM, device-timed (CUDA events),
               max over ranks
  e2e          the same through the public API with host buffers: grid.run(K) + the host->device upload
               of the waveform tables and the device->host read-back of every detector sample, wall clock
  roofline     the half-step kernel: algorithmic bytes per launch / its average duration (CUDA events,
               live) against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the oracle (CPU port of the reference algorithm, torch-CPU, all host threads) on a bounded
               sample of the same workload
--impl reference times that CPU port alone on the same config / metric / unit.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRID_SPACING = 77.5e-9
PML_CELLS = 10


def build_c4(fd, n, pml=PML_CELLS, balance=False):
    """configs[3]: six PMLs, centre PointSource, LineDetector along z through the centre."""
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    kw = {}
    if balance:
        # planes inside an x-PML move 13 instead of 9 words per cell and half-step: give their ranks fewer planes
        kw["x_plane_cost"] = [13.0 / 9.0 if (i < pml or i >= nx - pml) else 1.0 for i in range(nx)]
    g = fd.Grid(shape=(nx, ny, nz), grid_spacing=GRID_SPACING, **kw)
    g[0:pml, :, :] = fd.PML()
    g[-pml:, :, :] = fd.PML()
    g[:, 0:pml, :] = fd.PML()
    g[:, -pml:, :] = fd.PML()
    g[:, :, 0:pml] = fd.PML()
    g[:, :, -pml:] = fd.PML()
    g[nx // 2, ny // 2, nz // 2] = fd.PointSource(period=20, name="src")
    g[nx // 2 + 4, ny // 2, pml + 2:nz - pml - 2] = fd.LineDetector(name="line")
    return g


def algorithmic_bytes_per_cell_step(n, w, pml=PML_CELLS):
    """w*(18 + 8*M/N); M/N = PML slab memberships per cell = 2*pml*(1/Nx+1/Ny+1/Nz)
    (= 6*pml/n for a cube)."""
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    return w * (18.0 + 8.0 * 2 * pml * (1.0 / nx + 1.0 / ny + 1.0 / nz))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every 5 ms; the
    recipe's nvidia-smi line polls too slowly for a 40 ms region)."""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.thread = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates all GPUs of the box: map through CUDA_VISIBLE_DEVICES if set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.replace(",", "").isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def poll():
                while not self.stop:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.rows.append((sm, reasons))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.pynvml = pynvml
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.thread is not None:
            self.thread.join(timeout=1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.pynvml
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = []
        for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                           ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                           ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            flag = getattr(nv, attr, None) or getattr(nv, attr.replace("Event", "Throttle"), 0)
            if flag and (bits & flag):
                names.append(name)
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(self.max_sm), "reasons": names,
                "samples": len(sm)}


def cpu_port_rate(n, steps, dtype, threads=None):
    """Mcell-updates/s of the oracle (CPU port of the reference's algorithm) on an n^3 sample of the
    workload, torch-CPU with all host threads."""
    import torch
    from oracle import yee_oracle as yo
    if threads:
        torch.set_num_threads(threads)
    yo.set_backend("torch", dtype)
    try:
        g = build_c4(yo, n)
        g.run(1)                                   # warm-up step (allocations, first-touch)
        t0 = time.perf_counter()
        g.run(steps)
        dt = time.perf_counter() - t0
    finally:
        yo.set_backend("numpy", "float64")
    return n ** 3 * steps / dt / 1e6, dt, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference algorithm's CPU port on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_size
    steps = max(1, args.steps)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_port_rate(n, 1, args.dtype)
    rate, dt, threads = cpu_port_rate(n, steps, args.dtype)
    sample = f"{n}^3 sample of the {args.size}^3 workload, {steps} steps, torch-CPU {args.dtype}"
    line = {
        "impl": "reference", "metric": "Mcell-updates/s (3D Yee E+H step)", "value": rate,
        "unit": "Mcell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if args.dtype == "float32" else "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": rate, "unit": "Mcell-updates/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": rate, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def shape_of(args):
    if getattr(args, "shape", None):
        return tuple(int(v) for v in args.shape.split(","))
    return (args.size, args.size, args.size)


def workload_config(args):
    shape = shape_of(args)
    return {"workload": f"BASELINE configs[3]: 3D {shape[0]}x{shape[1]}x{shape[2]} {args.dtype} Yee grid, {PML_CELLS}-cell PML "
                        f"on all six faces, PointSource(period=20) at centre, LineDetector; x-slab sharded, halo exchange per half-step",
            "grid": list(shape), "pml_cells": PML_CELLS, "parallelism": f"x-slabs x{args.gpus}",
            "l2_policy": "inputs larger than L2 (fields are 24 GiB at 1024^3; every step streams all of them)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--cpu-size", type=int, default=320, help="edge of the CPU-baseline sample grid")
    ap.add_argument("--cpu-steps", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--x-chunk", type=int, default=0)
    ap.add_argument("--no-balance", action="store_true", help="equal plane counts per rank instead of equal cost")
    ap.add_argument("--shape", default=None, help="nx,ny,nz instead of --size (experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    import fdtd_b200 as fd
    from fdtd_b200 import _capi
    fd.set_backend("cuda." + args.dtype)
    lib = _capi.load()
    n, K, W = shape_of(args), args.steps, args.warmup
    cells = n[0] * n[1] * n[2]
    w = 4 if args.dtype == "float32" else 8

    grid = build_c4(fd, n, balance=not args.no_balance)
    grid._x_chunk = args.x_chunk
    det = grid.detectors[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up --------------------------------------------------------------------------------
    grid.run(W, progress_bar=False)
    eng = grid._engine
    eng.flush_detectors()
    barrier()

    # ---- value: K steps, fields resident, device-timed ---------------------------------------------
    eng._ensure_wave(grid.time_steps_passed, K)
    launches0 = lib.fdtd_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record()
        grid.run(K, progress_bar=False)
        stop.record()
        barrier()
    launches = lib.fdtd_launch_count() - launches0
    ms = torch.tensor([start.elapsed_time(stop)], device="cuda")
    per_rank_ms = [float(ms.item())]
    if world > 1:
        gathered = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(gathered, ms)
        per_rank_ms = [float(t.item()) for t in gathered]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = cells * K / (ms * 1e-3) / 1e6
    eng.flush_detectors()

    # ---- e2e: public API, host buffers in the timed region -----------------------------------------
    n_det_before = len(det.E)
    eng._wave = None                       # the waveform table is rebuilt and uploaded inside the region
    barrier()
    t0 = time.perf_counter()
    grid.run(K, progress_bar=False)
    traces = (det.E, det.H)                # flushes the device ring to the host
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = cells * K / float(e2e_s.item()) / 1e6
    assert len(traces[0]) == n_det_before + K
    h2d = sum(1 for _ in grid.sources) * eng._wave[1] * w / K
    d2h = 2 * det._n_points * 3 * w

    # ---- roofline: the half-step kernel alone, live ------------------------------------------------
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = eng.desc
    reps = max(4, min(K, 10))
    q_now = grid.time_steps_passed
    eng._ensure_wave(q_now, 1)
    barrier()
    eng.quiesce()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(reps):
        _capi.check(lib, lib.fdtd_e_halfstep(C.byref(d), 0, d.Nx, q_now, 0, st))
        _capi.check(lib, lib.fdtd_h_halfstep(C.byref(d), 0, d.Nx, q_now, 0, st))
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / (2 * reps)
    cells_local = d.Nx * n[1] * n[2]
    bytes_per_launch = algorithmic_bytes_per_cell_step(n, w) / 2 * cells_local
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath))
            if list(t.get("size", [])) == list(n) and t.get("dtype") == args.dtype and world == 1:
                traffic = t.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "fdtd::halfstep_kernel (E and H half-steps)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel_ms_per_launch": kernel_ms,
                "bytes_per_cell_step": algorithmic_bytes_per_cell_step(n, w)}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, dt, threads = cpu_port_rate(args.cpu_size, args.cpu_steps, args.dtype)
        cpu = {"value": rate, "unit": "Mcell-updates/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_size}^3 sample of the workload, {args.cpu_steps} steps, oracle on torch-CPU "
                         f"{args.dtype} ({dt:.1f} s)"}

    if rank == 0:
        line = {
            "metric": "Mcell-updates/s (3D Yee E+H step)", "value": value, "unit": "Mcell-updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if w == 4 else "f64", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "Mcell-updates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
            "per_rank_ms_per_step": [round(t / K, 4) for t in per_rank_ms],
            "hbm_roofline_frac_whole_step": (algorithmic_bytes_per_cell_step(n, w) * cells * K
                                             / (ms * 1e-3) / 1e9) / (peak * world),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
