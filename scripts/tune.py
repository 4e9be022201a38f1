"""Kernel tuning harness (development tool).

  python scripts/tune.py build      # here, no GPU: nvcc every variant into fdtd_b200/_variants/
  python scripts/tune.py run        # on the B200: time the E+H half-step kernels of every variant

Each variant is the same source compiled with different -D knobs (yee_kernels.cuh); timing is the
bench's kernel-only loop (CUDA events around 2*reps half-step launches at size^3 float32, six PMLs).
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "fdtd_b200", "_variants")

VARIANTS = {
    "default": [],
    # register-tiled fused E+H kernel (yee_fused_eh.cuh, FDTD_B200_FUSE_EH=2): rows per thread / warps per block /
    # blocks per SM -- time with scripts/gpu_fused_rt.sh
    # cp.async-pipelined fused kernel (FDTD_B200_FUSE_EH=3): tile and blocks per SM (shared memory: 3 stages + 2 tiles)
    "pipe_r4l32_mb3": ["-DFDTD_FUSED_ROWS=4", "-DFDTD_FUSED_LANES=32", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=3"],
    "pipe_r4l32_mb2": ["-DFDTD_FUSED_ROWS=4", "-DFDTD_FUSED_LANES=32", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=2"],
    "pipe_r8l32_mb1": ["-DFDTD_FUSED_ROWS=8", "-DFDTD_FUSED_LANES=32", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=1"],
    "pipe_r8l16_mb2": ["-DFDTD_FUSED_ROWS=8", "-DFDTD_FUSED_LANES=16", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=2"],
    "pipe_r2l32_mb4": ["-DFDTD_FUSED_ROWS=2", "-DFDTD_FUSED_LANES=32", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=4"],
    "pipe_r4l16_mb4": ["-DFDTD_FUSED_ROWS=4", "-DFDTD_FUSED_LANES=16", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=4"],
    # 31 core lanes + 1 halo lane = one full warp per tile row (5 or 8 warps per block, none partially filled)
    "pipe_r4l31_mb3": ["-DFDTD_FUSED_ROWS=4", "-DFDTD_FUSED_LANES=31", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=3"],
    "pipe_r4l31_mb2": ["-DFDTD_FUSED_ROWS=4", "-DFDTD_FUSED_LANES=31", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=2"],
    "pipe_r7l31_mb2": ["-DFDTD_FUSED_ROWS=7", "-DFDTD_FUSED_LANES=31", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=2"],
    # session 13: z+1 neighbour by shuffle, psi of the block's z slab staged by bulk copies
    "pipe_v2_noshfl": ["-DFDTD_FUSED_SHFL=0"],
    "pipe_v2_nopsi": ["-DFDTD_FUSED_PSI_STAGE=0"],
    "pipe_v2_old": ["-DFDTD_FUSED_SHFL=0", "-DFDTD_FUSED_PSI_STAGE=0"],
    "pipe_v2_r8": ["-DFDTD_FUSED_ROWS=8"],
    # session 17: psi of x-slab planes loaded at the top of the iteration
    "pipe_v5_noearly": ["-DFDTD_FUSED_EARLY_XPSI=0"],
    "pipe_tma_r15l31_mb1": ["-DFDTD_FUSED_ROWS=15", "-DFDTD_FUSED_LANES=31", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=1"],
    "pipe_tma_r3l31_mb4": ["-DFDTD_FUSED_ROWS=3", "-DFDTD_FUSED_LANES=31", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=4"],
    "pipe_r15l31_mb1": ["-DFDTD_FUSED_ROWS=15", "-DFDTD_FUSED_LANES=31", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=1"],
    "pipe_r3l31_mb4": ["-DFDTD_FUSED_ROWS=3", "-DFDTD_FUSED_LANES=31", "-DFDTD_FUSED_PIPE_MIN_BLOCKS=4"],
    "mat_mb2": ["-DFDTD_MAT_MIN_BLOCKS=2"],
    "mat_mb2_noinl": ["-DFDTD_MAT_MIN_BLOCKS=2", "-DFDTD_SPECIAL_NOINLINE=1"],
    "mat_noinl": ["-DFDTD_SPECIAL_NOINLINE=1"],
    "pfcap": ["-DFDTD_PREFETCH_CAP=1"],
    "pfcap_pf2": ["-DFDTD_PREFETCH_CAP=1", "-DFDTD_PREFETCH_PLANES=2"],
    "nopf": ["-DFDTD_PREFETCH_PLANES=0"],
    "post_inline": ["-DFDTD_POST_INLINE=1"],
    "hdown": ["-DFDTD_H_DOWNWARD=1"],
    "special_noinline": ["-DFDTD_SPECIAL_NOINLINE=1"],
    "lanes16": ["-DFDTD_MAX_LANES_Z=16"],
    "lanes8": ["-DFDTD_MAX_LANES_Z=8"],
    "lanes16_pf2": ["-DFDTD_MAX_LANES_Z=16", "-DFDTD_PREFETCH_PLANES=2"],
    "noinl": ["-DFDTD_NOINLINE_SLABS=1"],
    "noinl_mb4": ["-DFDTD_NOINLINE_SLABS=1", "-DFDTD_MIN_BLOCKS=4"],
    "pfG": ["-DFDTD_PREFETCH_WHAT=1"],
    "pfF": ["-DFDTD_PREFETCH_WHAT=2"],
    "cs": ["-DFDTD_STREAM_HINTS=1"],
    "noinl_cs": ["-DFDTD_NOINLINE_SLABS=1", "-DFDTD_STREAM_HINTS=1"],
    "base": ["-DFDTD_MIN_BLOCKS=2", "-DFDTD_PREFETCH_PLANES=0"],
    "mb3": ["-DFDTD_MIN_BLOCKS=3"],
    "mb4": ["-DFDTD_MIN_BLOCKS=4"],
    "pf1": ["-DFDTD_PREFETCH_PLANES=1"],
    "pf2": ["-DFDTD_PREFETCH_PLANES=2"],
    "pf4": ["-DFDTD_PREFETCH_PLANES=4"],
    "mb3_pf2": ["-DFDTD_MIN_BLOCKS=3", "-DFDTD_PREFETCH_PLANES=2"],
    "mb3_pf1": ["-DFDTD_MIN_BLOCKS=3", "-DFDTD_PREFETCH_PLANES=1"],
    "mb3_pf3": ["-DFDTD_MIN_BLOCKS=3", "-DFDTD_PREFETCH_PLANES=3"],
    "mb4_pf2": ["-DFDTD_MIN_BLOCKS=4", "-DFDTD_PREFETCH_PLANES=2"],
    "t128_mb6_pf2": ["-DFDTD_BLOCK_THREADS=128", "-DFDTD_MIN_BLOCKS=6", "-DFDTD_PREFETCH_PLANES=2"],
    "t128_mb4_pf2": ["-DFDTD_BLOCK_THREADS=128", "-DFDTD_MIN_BLOCKS=4", "-DFDTD_PREFETCH_PLANES=2"],
    "v2_mb4_pf2": ["-DFDTD_MAX_VEC_F32=2", "-DFDTD_MIN_BLOCKS=4", "-DFDTD_PREFETCH_PLANES=2"],
    "v2_mb6_pf2": ["-DFDTD_MAX_VEC_F32=2", "-DFDTD_MIN_BLOCKS=6", "-DFDTD_PREFETCH_PLANES=2"],
}


def build():
    import __graft_entry__ as entry
    os.makedirs(VDIR, exist_ok=True)
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else list(VARIANTS)
    for name, defs in VARIANTS.items():
        if name not in only:
            continue
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = ["nvcc"] + entry.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-I", os.path.join(ROOT, "include"),
                                                  "-I", os.path.join(ROOT, "fdtd_b200", "csrc"), entry.SRC, "-o", out]
        r = subprocess.run(cmd, check=True, capture_output=True, text=True)
        lines = r.stderr.splitlines()
        for n, l in enumerate(lines):
            if "Compiling entry function" in l and "fused_eh_rt_kernelIfLi4" in l and name.startswith("rt_"):
                print(name, "FUSED-RT", " | ".join(x.strip() for x in lines[n + 1:n + 4]))
            elif "Compiling entry function" in l and "fused_eh_pipe_kernelIfLi4" in l and name.startswith("pipe"):
                print(name, "PIPE", " | ".join(x.strip() for x in lines[n + 1:n + 4]))
            elif "Compiling entry function" in l and ("fused_eh_kernelIfLi4" in l) and name.startswith("fz"):
                print(name, "FUSED", " | ".join(x.strip() for x in lines[n + 1:n + 4]))
            elif "Compiling entry function" in l and not name.startswith("fz") and ("halfstep_kernelIfLi4ELb" in l or ("MAX_VEC_F32=2" in " ".join(defs) and "halfstep_kernelIfLi2ELb" in l)):
                print(name, "E" if "ELb1" in l else "H", " | ".join(x.strip() for x in lines[n + 1:n + 4]))


def run(size=1024, dtype="float32", reps=10, chunks=(32,)):
    import torch
    import fdtd_b200 as fd
    from fdtd_b200 import _capi
    from fdtd_b200.backend import backend
    sys.path.insert(0, ROOT)
    from bench import build_c4, algorithmic_bytes_per_cell_step
    fd.set_backend("cuda." + dtype)
    w = 4 if dtype == "float32" else 8
    results = []
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(VARIANTS)
    for name in names:
        path = os.path.join(VDIR, f"lib_{name}.so")
        if not os.path.exists(path):
            print("missing", path)
            continue
        lib = _capi.bind(path)
        backend.lib = lib
        for chunk in chunks:
            grid = build_c4(fd, size)
            grid._x_chunk = chunk
            grid.run(2, progress_bar=False)
            eng = grid._engine
            d = eng.desc
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            out = {}
            for label, fn in (("E", lib.fdtd_e_halfstep), ("H", lib.fdtd_h_halfstep)):
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    _capi.check(lib, fn(C.byref(d), 0, d.Nx, grid.time_steps_passed, 0, st))
                b.record()
                torch.cuda.synchronize()
                out[label] = a.elapsed_time(b) / reps
            if isinstance(size, int):
                bytes_launch = algorithmic_bytes_per_cell_step(size, w) / 2 * size ** 3
            else:
                nx, ny, nz = size
                mem = 2 * 10 * (1.0 / nx + 1.0 / ny + 1.0 / nz)
                bytes_launch = w * (9.0 + 4.0 * mem) * nx * ny * nz
            gbs = bytes_launch / ((out["E"] + out["H"]) / 2 * 1e-3) / 1e9
            rec = {"variant": name, "x_chunk": chunk, "E_ms": round(out["E"], 3), "H_ms": round(out["H"], 3),
                   "GBs": round(gbs, 1)}
            print(json.dumps(rec), flush=True)
            results.append(rec)
            del grid, eng, d
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    return results


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        size = os.environ.get("TUNE_SIZE", "1024")
        size = int(size) if "x" not in size else tuple(int(v) for v in size.split("x"))
        chunks = tuple(int(c) for c in os.environ.get("TUNE_CHUNKS", "0").split(","))
        run(size=size, chunks=chunks, dtype=os.environ.get("TUNE_DTYPE", "float32"))
