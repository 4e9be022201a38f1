#!/bin/bash
# Round-2 GPU session 11 (development tool, 1 GPU): fused (TMA) against two half-steps by grid size and by slab
# thickness (x-extent) -- the thresholds of the automatic mode.
set -u
mkdir -p gpurun_out/r2_s11
python - <<'PY' 2>&1 | tee gpurun_out/r2_s11/fused_sizes_tma.log
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fdtd_b200 as fd
from bench import build_c4
fd.set_backend("cuda.float32")
def t(shape, fuse):
    g = build_c4(fd, shape); g._fuse_eh = fuse
    g.run(4, progress_bar=False); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.run(20, progress_bar=False); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    del g
    return ms
for n in (256, 384, 512, 640, 768, 896, 1024):
    a, b = t(n, 0), t(n, 1)
    print(f"{n}^3: two half-steps {a:.3f} ms, fused {b:.3f} ms ({100 * (a / b - 1):+.1f} %)", flush=True)
for nx in (64, 128, 192, 256, 512):
    a, b = t((nx, 1024, 1024), 0), t((nx, 1024, 1024), 1)
    print(f"{nx}x1024x1024: two half-steps {a:.3f} ms, fused {b:.3f} ms ({100 * (a / b - 1):+.1f} %)", flush=True)
fd.set_backend("cuda.float64")
for n in (256, 512, 640):
    a, b = t(n, 0), t(n, 1)
    print(f"float64 {n}^3: two half-steps {a:.3f} ms, fused {b:.3f} ms ({100 * (a / b - 1):+.1f} %)", flush=True)
PY
