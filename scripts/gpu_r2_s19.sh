#!/bin/bash
# Round-2 GPU session 19 (development tool, 1 GPU): x-chunk length sweep of the final fused kernel (1024^3 and a
# 256-plane slab), DRAM bytes at two lengths.
set -u
out=gpurun_out/r2_s19
mkdir -p $out
{
for xc in 24 28 32 36 40 44 49; do
  echo "# x_chunk=$xc"; X_CHUNK=$xc python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
} | tee $out/xchunk_1024.log
python - <<'PY' 2>&1 | tee $out/xchunk_slabs.log
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
for nx in (128, 256):
    print(f"{nx}x1024x1024:", ", ".join(f"x_chunk {xc}: {t((nx, 1024, 1024), xc):.4f}" for xc in (0, 22, 26, 32, 37, 43, 52)), flush=True)
PY
for xc in 40 49; do
X_CHUNK=$xc ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fused_eh_pipe -s 6 -c 1 --csv --log-file $out/metrics_xc$xc.csv python scripts/bench_configs.py c4 > /dev/null 2>&1
echo "x_chunk=$xc"; grep fused $out/metrics_xc$xc.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done | tee $out/dram_by_xchunk.log
