#!/bin/bash
# Round-2 GPU session 14 (development tool, 1 GPU): the fused kernel with the split block barrier, the CPML tables in
# shared memory and the wave-aware chunking -- correctness, sanitizers, timing, thresholds of the automatic mode.
set -u
out=gpurun_out/r2_s14
mkdir -p $out /tmp/rep
python -m pytest tests/test_gpu_parity.py -x -q -k "fused" 2>&1 | tail -4 | tee $out/pytest_fused.log
for tool in racecheck memcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_scenes.py only-fused > $out/$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|error" $out/$tool.log | tail -5
done
{
echo "# default build (split barrier, tables in shared memory)"; python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
for lib in fdtd_b200/_variants/lib_pipe_v3_*.so; do
  echo "# $lib"; TUNE_LIB=$lib python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
} | tee $out/variants.log
python - <<'PY' 2>&1 | tee $out/fused_sizes.log
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
for n in (384, 512, 576, 640, 768, 1024):
    a, b = t(n, 0), t(n, 1)
    print(f"{n}^3: two half-steps {a:.3f} ms, fused {b:.3f} ms ({100 * (a / b - 1):+.1f} %)", flush=True)
for nx in (64, 96, 128, 256, 512):
    a, b = t((nx, 1024, 1024), 0), t((nx, 1024, 1024), 1)
    print(f"{nx}x1024x1024: two half-steps {a:.3f} ms, fused {b:.3f} ms ({100 * (a / b - 1):+.1f} %)", flush=True)
fd.set_backend("cuda.float64")
for n in (256, 384, 512):
    a, b = t(n, 0), t(n, 1)
    print(f"float64 {n}^3: two half-steps {a:.3f} ms, fused {b:.3f} ms ({100 * (a / b - 1):+.1f} %)", flush=True)
PY
ncu --set full --clock-control none --import-source on -k regex:fused_eh_pipe -s 6 -c 1 -o /tmp/rep/c4_fused_v3 python scripts/bench_configs.py c4 > $out/c4_fused_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/rep/c4_fused_v3.ncu-rep > $out/ncu_summary_fused_v3.txt 2>&1
cp /tmp/rep/c4_fused_v3.ncu-rep $out/
tail -8 $out/ncu_summary_fused_v3.txt
