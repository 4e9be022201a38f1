#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of the fused 3-D Yee step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1|c2|c3|c4|c5]

Workloads (`--config`, BASELINE.json configs[0..4]; inputs are SURVEY.md section 8d's, all synthetic and closed-form):
  c1  161x97x1 float64 quick-start grid: four 10-cell PMLs, LineSource, Object n=1.7, LineDetector (launch-bound)
  c2  256^3 float64, six 10-cell PMLs, PointSource + BlockDetector
  c3  512^3 float32, six PMLs, PlaneSource, AbsorbingObject + anisotropic lens, LineDetector
  c4  1024^3 float32, six PMLs, PointSource + LineDetector                      <- default, the headline
  c5  (256 N)x1024x1024 float32 waveguide: x-PMLs, periodic y / z, GRIN Object over the middle half, PlaneSource
      (weak scaling: 256 x-planes per GPU; N = 8 is BASELINE's 2048x1024x1024)
For N > 1 the grid is split into x-slabs, one process per GPU (torchrun); ghost planes are stored straight into the
neighbour's memory over NVLink by the kernel that computes them, ordered by flag words -- no data-path collective.
A "step" is one full E+H update of the whole grid including PML, material, source and detector work.

Printed JSON (one line, rank 0):
  value        Mcell-updates/s over all GPUs, fields resident in HBM, device-timed (CUDA events), max over ranks
  e2e          the same through the public API with host buffers: grid.run(K) + the host->device upload of the waveform
               tables and the device->host read-back of every detector sample, wall clock
  roofline     the half-step kernel: algorithmic bytes per launch / its average duration (CUDA events, live) against the
               measured HBM copy bandwidth in MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch from the committed
               ncu capture of this kernel on this config (profiles/traffic.json), null when there is none
  cpu_baseline the oracle (CPU port of the reference algorithm) on a bounded sample of the same workload: torch-CPU with
               all host threads, plus `numpy_f64` = the reference's default backend arithmetic (numpy float64, 1 thread)
  gpu_eager_baseline  the oracle's torch flavour on device="cuda": the same slicing code as eager ATen kernels, which is
               how the reference's own `torch.cuda` backends run on a GPU (fdtd/backend.py:322-355)
  parity       N > 1: a reduced scene of the same structure run x-sharded and unsharded outside the timed region,
               compared bit for bit
--impl reference times the CPU port alone on the same config / metric / unit (rank 0 only under torchrun).
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


# ------------------------------------------------------------------------------------------------ workloads
def six_pml(fd, g, t=PML_CELLS):
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, -t:, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[:, :, -t:] = fd.PML()


# what a plane inside an x-PML costs relative to an ordinary one, measured on the B200 (profiles/r2_s16, r2_s17): the two
# half-steps 1.56 (13 instead of 9 words per cell and half-step), the single-pass E+H kernel more (20 instead of 12 words
# per cell and step, and psi arrives by plain loads instead of staged copies)
XPML_PLANE_COST, XPML_PLANE_COST_FUSED = 1.56, 1.93


def pml_plane_cost(nx, pml=PML_CELLS, lo=True, hi=True, weight=XPML_PLANE_COST):
    """planes inside an x-PML cost more than ordinary ones: give their ranks fewer planes"""
    return [weight if ((lo and i < pml) or (hi and i >= nx - pml)) else 1.0 for i in range(nx)]


def fused_steps_expected(fd, shape, itemsize=4):
    """will the ranks of an x-sharded homogeneous grid run the single-pass E+H kernel? (fdtd_b200/engine.py's test)"""
    import torch.distributed as dist
    from fdtd_b200 import engine
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    nx, ny, nz = shape
    tile_z = 31 * (16 // itemsize)
    return (os.environ.get("FDTD_B200_FUSE_EH", "2") != "0" and ny * nz * itemsize >= engine.FUSE_EH_MIN_PLANE_BYTES
            and nz >= engine.FUSE_EH_MIN_Z_FILL * (-(-nz // tile_z) * tile_z)
            and nx // world >= (engine.FUSE_EH_MIN_SLAB if world > 1 else engine.FUSE_EH_MIN_PLANES))


def build_c4(fd, n, pml=PML_CELLS, balance=False, **kw):
    """configs[3]: six PMLs, centre PointSource, LineDetector along z through the centre."""
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    if balance:
        fused = fused_steps_expected(fd, (nx, ny, nz))
        kw["x_plane_cost"] = pml_plane_cost(nx, pml, weight=XPML_PLANE_COST_FUSED if fused else XPML_PLANE_COST)
    g = fd.Grid(shape=(nx, ny, nz), grid_spacing=GRID_SPACING, **kw)
    six_pml(fd, g, pml)
    g[nx // 2, ny // 2, nz // 2] = fd.PointSource(period=20, name="src")
    g[nx // 2 + 4, ny // 2, pml + 2:nz - pml - 2] = fd.LineDetector(name="line")
    return g


def build_c1(fd, balance=False, **kw):
    """configs[0]: the README quick-start (fdtd README.md:137-297)."""
    g = fd.Grid(shape=(25e-6, 15e-6, 1), grid_spacing=155e-9, **kw)
    g[11:32, 30:84, 0] = fd.Object(permittivity=1.7 ** 2, name="object")
    g[7.5e-6:8.0e-6, 11.8e-6:13.0e-6, 0] = fd.LineSource(period=1550e-9 / (3e8), name="source")
    g[12e-6, :, 0] = fd.LineDetector(name="detector")
    g[0:10, :, :] = fd.PML()
    g[-10:, :, :] = fd.PML()
    g[:, 0:10, :] = fd.PML()
    g[:, -10:, :] = fd.PML()
    g[:, :, 0] = fd.PeriodicBoundary(name="zbounds")
    return g


def build_c2(fd, n=256, balance=False, **kw):
    """configs[1]: float64, six PMLs, PointSource + BlockDetector (3x3x3 points)."""
    if balance:
        kw["x_plane_cost"] = pml_plane_cost(n)
    g = fd.Grid(shape=(n, n, n), grid_spacing=GRID_SPACING, **kw)
    six_pml(fd, g)
    c = n // 2
    g[c, c, c] = fd.PointSource(period=20)
    g[c + 10:c + 12, c:c + 2, c:c + 2] = fd.BlockDetector()
    return g


def build_c3(fd, n=512, balance=False, **kw):
    """configs[2]: float32, six PMLs, PlaneSource, AbsorbingObject slab and an anisotropic plano-convex lens
    (closed-form mask), LineDetector along x through the focus.  All positions scale with n / 512."""
    import numpy as np
    s = n / 512.0
    r = lambda v: int(round(v * s))
    if balance:
        kw["x_plane_cost"] = pml_plane_cost(n)
    g = fd.Grid(shape=(n, n, n), grid_spacing=GRID_SPACING, **kw)
    six_pml(fd, g, PML_CELLS if n >= 128 else 4)
    g[r(60), :, :] = fd.PlaneSource(period=20, polarization="z")
    g[r(150):r(200), r(100):r(412), r(100):r(412)] = fd.AbsorbingObject(permittivity=2.5, conductivity=1.5e4)
    x0, x1, a, b = r(260), r(324), r(128), r(384)
    i = np.arange(x0, x1)[:, None, None]
    j = np.arange(a, b)[None, :, None]
    k = np.arange(a, b)[None, None, :]
    mask = ((j - n // 2) ** 2 + (k - n // 2) ** 2 + (i - r(164)) ** 2) <= r(160) ** 2
    P = np.ones((x1 - x0, b - a, b - a, 3))
    P[mask] = (2.25, 2.25, 2.5)
    g[x0:x1, a:b, a:b] = fd.AnisotropicObject(permittivity=P)
    g[r(20):r(492), n // 2, n // 2] = fd.LineDetector()
    return g


def build_c5(fd, nx, n=1024, balance=False, **kw):
    """configs[4]: waveguide, PML on x, periodic y and z, GRIN medium (linear ramp along y) over the middle half,
    PlaneSource, LineDetector along x."""
    import numpy as np
    t = PML_CELLS if nx >= 64 else 3
    if balance:
        # words per cell-step of each x-plane: 18, + 8 inside an x-PML (psi), + 3 where the GRIN eps^-1 is streamed
        kw["x_plane_cost"] = [(18.0 + 8.0 * (i < t or i >= nx - t) + 3.0 * (nx // 4 <= i < 3 * nx // 4)) / 18.0
                              for i in range(nx)]
    g = fd.Grid(shape=(nx, n, n), grid_spacing=GRID_SPACING, **kw)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0, :] = fd.PeriodicBoundary()
    g[:, :, 0] = fd.PeriodicBoundary()
    ramp = (1.0 + 1.25 * np.arange(n) / (n - 1.0)).reshape(1, n, 1)
    g[nx // 4:3 * nx // 4, :, :] = fd.Object(permittivity=ramp)
    g[min(100, nx // 3), :, :] = fd.PlaneSource(period=20, polarization="z")
    g[2 * t:nx - 2 * t, n // 2, n // 2] = fd.LineDetector()
    return g


class Workload:
    """one BASELINE config: builder, dtype, algorithmic words per cell-step, reduced twin for parity."""

    def __init__(self, name, world):
        self.name, self.world = name, world
        pml = 2.0 * PML_CELLS
        if name == "c1":
            self.dtype, self.shape, self.steps = "float64", (161, 97, 1), 1000
            self.words = 18 + 8 * pml * (1 / 161 + 1 / 97) + 3 * (21 * 54) / (161 * 97)
            self.build = lambda fd, **kw: build_c1(fd, **kw)
            self.small = lambda fd, **kw: build_c1(fd, **kw)
            self.text = ("BASELINE configs[0]: 2D quick-start grid 161x97x1 float64, 10-cell PMLs, LineSource, Object "
                         "n=1.7, LineDetector")
        elif name == "c2":
            self.dtype, self.shape, self.steps = "float64", (256, 256, 256), 2000
            self.words = 18 + 8 * pml * 3 / 256
            self.build = lambda fd, **kw: build_c2(fd, 256, **kw)
            self.small = lambda fd, **kw: build_c2(fd, 48, **kw)
            self.text = "BASELINE configs[1]: 3D 256^3 float64, PML on all six faces, PointSource + BlockDetector"
        elif name == "c3":
            self.dtype, self.shape, self.steps = "float32", (512, 512, 512), 500
            absorber, lens = 50 * 312 * 312 / 512 ** 3, 64 * 256 * 256 / 512 ** 3
            self.words = 18 + 8 * pml * 3 / 512 + 6 * absorber + 3 * lens
            self.build = lambda fd, **kw: build_c3(fd, 512, **kw)
            self.small = lambda fd, **kw: build_c3(fd, 64, **kw)
            self.text = ("BASELINE configs[2]: 3D 512^3 float32, AbsorbingObject + anisotropic (Nx,Ny,Nz,3) permittivity "
                         "lens, PML on all six faces, PlaneSource, LineDetector")
        elif name == "c4":
            self.dtype, self.shape, self.steps = "float32", (1024, 1024, 1024), 200
            self.words = 18 + 8 * pml * 3 / 1024
            self.build = lambda fd, **kw: build_c4(fd, 1024, **kw)
            self.small = lambda fd, **kw: build_c4(fd, (64, 48, 48), pml=6, **kw)
            self.text = ("BASELINE configs[3]: 3D 1024x1024x1024 float32 Yee grid, 10-cell PML on all six faces, "
                         "PointSource(period=20) at centre, LineDetector; x-slab sharded, halo exchange per half-step")
        elif name == "c5":
            nx = 256 * world
            self.dtype, self.shape, self.steps = "float32", (nx, 1024, 1024), 100
            self.words = 18 + 8 * pml / nx + 3 * 0.5
            self.build = lambda fd, **kw: build_c5(fd, nx, **kw)
            self.small = lambda fd, **kw: build_c5(fd, 16 * max(2, world), 24, **kw)
            self.text = (f"BASELINE configs[4]: 3D {nx}x1024x1024 float32 periodic-y/z + PML-x waveguide with GRIN "
                         f"medium, PlaneSource (weak scaling: 256 x-planes per GPU; 8 GPUs = 2048x1024x1024)")
        else:
            raise ValueError(name)
        self.scaling = "weak" if name == "c5" else "strong"
        self.w = 4 if self.dtype == "float32" else 8
        self.cells = self.shape[0] * self.shape[1] * self.shape[2]

    @property
    def bytes_per_cell_step(self):
        return self.w * self.words

    def config(self):
        return {"workload": self.text, "config": self.name, "grid": list(self.shape), "pml_cells": PML_CELLS,
                "parallelism": f"x-slabs x{self.world}",
                "l2_policy": ("inputs larger than L2: every step streams all fields "
                              f"({2 * 3 * self.cells * self.w / 2 ** 30:.1f} GiB) through HBM" if self.cells * self.w * 6 > (256 << 20)
                              else "working set fits in L2 (BASELINE config 0 is launch-bound, reported as is)")}


def algorithmic_bytes_per_cell_step(n, w, pml=PML_CELLS):
    """SURVEY.md section 8d for the c4 structure at any size: w*(18 + 8*M/N), M/N = 2*pml*(1/Nx+1/Ny+1/Nz)."""
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


# ------------------------------------------------------------------------------------ baselines (oracle port)
def sample_builder(wl, edge):
    """the workload's structure on a bounded grid the CPU finishes in seconds"""
    if wl.name == "c1":
        return (lambda fd: build_c1(fd)), 161 * 97, "the full 161x97x1 grid"
    if wl.name == "c2":
        return (lambda fd: build_c2(fd, edge)), edge ** 3, f"{edge}^3 sample of the 256^3 workload"
    if wl.name == "c3":
        return (lambda fd: build_c3(fd, edge)), edge ** 3, f"{edge}^3 sample of the 512^3 workload"
    if wl.name == "c5":
        return (lambda fd: build_c5(fd, edge, edge)), edge ** 3, f"{edge}^3 sample of the {wl.shape[0]}x1024x1024 workload"
    return (lambda fd: build_c4(fd, edge)), edge ** 3, f"{edge}^3 sample of the 1024^3 workload"


def oracle_rate(wl, edge, steps, kind="torch", dtype=None, device=None, threads=None):
    """Mcell-updates/s of the oracle (CPU port of the reference's algorithm) on a sample of the workload."""
    import torch
    from oracle import yee_oracle as yo
    if kind == "torch" and device is None:
        # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        torch.set_num_threads(threads or os.cpu_count() or 1)
    build, cells, what = sample_builder(wl, edge)
    yo.set_backend(kind, dtype or wl.dtype, device)
    sync = (lambda: torch.cuda.synchronize()) if device else (lambda: None)
    try:
        g = build(yo)
        g.run(1)                                   # warm-up step (allocations, first-touch)
        sync()
        t0 = time.perf_counter()
        g.run(steps)
        sync()
        dt = time.perf_counter() - t0
    finally:
        yo.set_backend("numpy", "float64")
    return cells * steps / dt / 1e6, dt, what, (torch.get_num_threads() if kind == "torch" and device is None else 1)


def run_reference(args, wl):
    """--impl reference: the reference algorithm's CPU port on this box's host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps = max(1, args.steps)
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_rate(wl, args.cpu_size, 1)
    rate, dt, what, threads = oracle_rate(wl, args.cpu_size, steps)
    line = {
        "impl": "reference", "metric": "Mcell-updates/s (3D Yee E+H step)", "value": rate,
        "unit": "Mcell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": wl.scaling,
        "vs_baseline": None, "dtype": "f32" if wl.dtype == "float32" else "f64", "data": "synthetic",
        "config": wl.config(),
        "cpu_baseline": {"value": rate, "unit": "Mcell-updates/s", "cores": threads, "kind": "port",
                         "sample": f"{what}, {steps} steps, oracle on torch-CPU {wl.dtype}"},
        "e2e": {"value": rate, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- parity at N > 1
def sharded_parity(fd, wl, world, rank, steps=24):
    """a reduced scene of the workload's structure, x-sharded over the job's ranks and (rank 0) unsharded, outside
    any timed region: final E, H and every detector trace must agree bit for bit."""
    import numpy as np
    import torch
    import scenes
    g = wl.small(fd)
    assert g._part.sharded and g._part.world == world
    g.run(steps // 2, progress_bar=False)
    for _ in range(steps - steps // 2):
        g.step()
    got = scenes.dump(g)                   # collective: gathers the slabs and the detector samples
    halo = "p2p" if g._engine._p2p else "nccl"
    ok, worst = True, 0.0
    if rank == 0:
        s = wl.small(fd, shard=False)
        s.run(steps, progress_bar=False)
        want = scenes.dump(s)
        for k in want:
            same = got[k].shape == want[k].shape and np.array_equal(got[k], want[k])
            ok = ok and same
            if not same and got[k].shape == want[k].shape:
                worst = max(worst, scenes.rel_l2(got[k], want[k]))
        nonzero = float(np.abs(want["E"]).max()) > 0
        ok = ok and nonzero
        del s
    del g
    torch.cuda.synchronize()
    out = {"sharded_equals_single": bool(ok), "scene": f"{wl.name} structure on {list(got['E'].shape[:3])}, {steps} steps "
           f"(run + step), {world} ranks, halo {halo}", "arrays": len(got), "worst_rel_l2": worst}
    if wl.name == "c4":
        # the temporally fused steps the full-size slabs run (pairs of single-pass E+H kernels, boundary planes into
        # the neighbours' second buffers) against the unsharded two-half-step path
        import ctypes
        shape = (max(32, 8 * world), 40, 136)
        g = build_c4(fd, shape, pml=5)
        g._fuse_eh = 1
        g.run(steps + 1, progress_bar=False)
        eng = g._engine
        active = bool(eng._p2p) and eng.lib.fdtd_fuse_eh_sharded_active(ctypes.byref(eng.desc), ctypes.byref(eng._p2p.h)) == 1
        got = scenes.dump(g)
        same = True
        if rank == 0:
            s = build_c4(fd, shape, pml=5, shard=False)
            s._fuse_eh = 0
            s.run(steps + 1, progress_bar=False)
            want = scenes.dump(s)
            same = all(np.array_equal(got[k], want[k]) for k in want) and float(np.abs(want["E"]).max()) > 0
            del s
        del g
        torch.cuda.synchronize()
        out["fused_steps"] = {"active": bool(active), "sharded_fused_equals_single_two_pass": bool(same),
                              "scene": f"c4 structure on {list(shape)}, {steps + 1} steps"}
        out["sharded_equals_single"] = bool(out["sharded_equals_single"] and same)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--mode", default=None, choices=[None, "float32", "float64", "float32x"],
                    help="engine precision mode (default: the config's dtype; float32x = float32 state, float64 arithmetic)")
    ap.add_argument("--cpu-size", type=int, default=0, help="edge of the CPU-baseline sample grid (0: per config)")
    ap.add_argument("--cpu-steps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--x-chunk", type=int, default=0)
    ap.add_argument("--no-balance", action="store_true", help="equal plane counts per rank instead of equal cost")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = Workload(args.config, max(world, args.gpus) if args.impl == "reference" else world)
    if not args.cpu_size:
        args.cpu_size = {"c1": 0, "c2": 192, "c3": 256, "c4": 320, "c5": 256}[wl.name]
    if not args.cpu_steps:
        args.cpu_steps = 400 if wl.name == "c1" else (12 if wl.dtype == "float64" else 16)

    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    import fdtd_b200 as fd
    from fdtd_b200 import _capi
    mode = args.mode or wl.dtype
    fd.set_backend("cuda." + mode)
    lib = _capi.load()
    K, W = args.steps, args.warmup
    cells, w = wl.cells, wl.w

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of the sharded path on this hardware (outside the timed region) -------------------------
    parity = sharded_parity(fd, wl, world, rank) if world > 1 else None
    barrier()

    grid = wl.build(fd, balance=not args.no_balance)
    grid._x_chunk = args.x_chunk
    det = grid.detectors[0]

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
    wa = 8 if mode != "float32" else 4
    h2d = len({id(s) for _, s in eng._src_entries}) * eng._wave[1] * wa / K
    d2h = 2 * det._n_points * 3 * w

    # ---- roofline: the dominant kernel alone, live -------------------------------------------------
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = eng.desc
    reps = max(4, min(K, 10))
    q_now = grid.time_steps_passed
    eng._ensure_wave(q_now, 2 * reps)
    barrier()
    eng.quiesce()
    torch.cuda.synchronize()
    peak, peak_src = measured_peak()
    cells_local = d.Nx * wl.shape[1] * wl.shape[2]

    def timed(fn):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        fn()
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1)

    def halfsteps():
        for _ in range(reps):
            _capi.check(lib, lib.fdtd_e_halfstep(C.byref(d), 0, d.Nx, q_now, 0, st))
            _capi.check(lib, lib.fdtd_h_halfstep(C.byref(d), 0, d.Nx, q_now, 0, st))

    kernel_ms = timed(halfsteps) / (2 * reps)
    bytes_per_launch = wl.bytes_per_cell_step / 2 * cells_local
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    fused = world == 1 and lib.fdtd_fuse_eh_active(C.byref(d)) == 1
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            t = json.load(open(tpath)).get(f"{wl.name}:{mode}:{'fused' if fused else 'halfstep'}")
            if t:
                traffic, traffic_src = t.get("dram_bytes_per_launch"), t.get("source")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "fdtd::halfstep_kernel (E and H half-steps, averaged)", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms_per_launch": kernel_ms,
                "bytes_per_cell_step": wl.bytes_per_cell_step, "words_per_cell_step": wl.words}
    if fused:
        # grid.run() executes pairs of single-pass E+H steps (12 field words per cell-step instead of 18; the psi
        # words are the same): that kernel is the dominant one, scored against ITS algorithmic bytes -- and, beside
        # it, against the two-pass bytes SURVEY 8d counts, which a single pass may exceed
        fused_ms = timed(lambda: _capi.check(lib, lib.fdtd_run(C.byref(d), q_now, 2 * reps, 0, st))) / (2 * reps)
        words12 = wl.words - 6.0
        b12 = w * words12 * cells_local
        halfstep = dict(roofline)
        roofline = {"bound": "hbm", "kernel": "fdtd::fused_eh_pipe_kernel (E and H in one pass; one launch per step)",
                    "achieved": b12 / (fused_ms * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": b12 / (fused_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": b12, "kernel_ms_per_launch": fused_ms,
                    "bytes_per_cell_step": w * words12, "words_per_cell_step": words12,
                    "two_pass_equivalent": {"bytes_per_cell_step": wl.bytes_per_cell_step,
                                            "achieved": 2 * bytes_per_launch / (fused_ms * 1e-3) / 1e9,
                                            "frac": 2 * bytes_per_launch / (fused_ms * 1e-3) / 1e9 / peak},
                    "halfstep_kernels": {k: halfstep[k] for k in ("achieved", "frac", "kernel_ms_per_launch",
                                                                  "algorithmic_bytes_per_launch")}}

    # ---- baselines (rank 0, N=1 only) ---------------------------------------------------------------
    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, dt, what, threads = oracle_rate(wl, args.cpu_size, args.cpu_steps)
        cpu = {"value": rate, "unit": "Mcell-updates/s", "cores": threads, "kind": "port",
               "sample": f"{what}, {args.cpu_steps} steps, oracle on torch-CPU {wl.dtype} ({dt:.1f} s)"}
        # the reference's default backend is numpy float64 (fdtd/backend.py:363): its arithmetic, one host thread
        nsz = args.cpu_size if wl.name == "c1" else min(args.cpu_size, 160)
        nst = 200 if wl.name == "c1" else 4
        rate, dt, what, _ = oracle_rate(wl, nsz, nst, kind="numpy", dtype="float64")
        cpu["numpy_f64"] = {"value": rate, "unit": "Mcell-updates/s", "cores": 1,
                            "sample": f"{what}, {nst} steps, oracle on numpy float64 ({dt:.1f} s)"}
    if rank == 0 and world == 1 and not args.no_gpu_eager_baseline:
        del grid, eng, d, det, traces
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        esz = 0 if wl.name == "c1" else (256 if wl.dtype == "float64" else 384)
        est = 200 if wl.name == "c1" else 10
        try:
            rate, dt, what, _ = oracle_rate(wl, esz, est, kind="torch", device="cuda")
            eager = {"value": rate, "unit": "Mcell-updates/s", "kind": "port on torch.cuda (eager ATen kernels)",
                     "sample": f"{what}, {est} steps, oracle torch flavour on device=cuda {wl.dtype} ({dt:.2f} s)"}
        except Exception as exc:             # never let the side measurement take the bench line with it
            eager = {"value": None, "error": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": "Mcell-updates/s (3D Yee E+H step)", "value": value, "unit": "Mcell-updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None,
            "dtype": {"float32": "f32", "float64": "f64", "float32x": "f32 storage / f64 arithmetic"}[mode],
            "data": "synthetic", "config": wl.config(),
            "e2e": {"value": e2e_value, "unit": "Mcell-updates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "parity": parity,
            "clocks": clocks.summary(),
            "per_rank_ms_per_step": [round(t / K, 4) for t in per_rank_ms],
            "hbm_roofline_frac_whole_step": (wl.bytes_per_cell_step * cells * K / (ms * 1e-3) / 1e9) / (peak * world),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
