#!/bin/bash
# Round-2 2-GPU session 20 (development tool; gpurun --gpus 2): x-chunks with one short chunk launched last --
# correctness (fused GPU tests, sharded fused tests), 1024^3 and slab timings, thin slabs on two GPUs, bench N=2.
set -u
out=gpurun_out/r2_s20
mkdir -p $out
python -m pytest tests/test_gpu_parity.py -x -q -k "fused" 2>&1 | tail -3 | tee $out/pytest_fused.log
python -m pytest tests/test_gpu_sharded.py -x -q -k "fused or c4small-float32-p2p" 2>&1 | tail -3 | tee $out/pytest_sharded.log
{
echo "# default"; python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
echo "# x_chunk=44"; X_CHUNK=44 python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
} | tee $out/c4.log
python - <<'PY' 2>&1 | tee $out/slabs.log
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fdtd_b200 as fd
from bench import build_c4
fd.set_backend("cuda.float32")
def t(shape, xc):
    g = build_c4(fd, shape); g._fuse_eh = 1; g._x_chunk = xc
    g.run(4, progress_bar=False); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.run(40, progress_bar=False); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 40
    del g
    return ms
for nx in (128, 256, 512):
    print(f"{nx}x1024x1024:", ", ".join(f"x_chunk {xc}: {t((nx, 1024, 1024), xc):.4f}" for xc in (0, 22, 32)), flush=True)
PY
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
{
$TR scripts/slab_bench.py 256 40 2>/dev/null | tail -1
$TR scripts/slab_bench.py 512 40 2>/dev/null | tail -1
} | tee $out/slab_bench.log
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $out/bench_c4_n2.json
cut -c1-330 $out/bench_c4_n2.json
