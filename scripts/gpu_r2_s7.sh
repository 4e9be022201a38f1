#!/bin/bash
# Round-2 GPU session 7 (development tool, 1 GPU): fused default path, material-kernel variants on c3 and c5slab, fused at 512^3.. 768^3
set -u
out=gpurun_out/r2_s7
mkdir -p $out
python -m pytest tests/test_gpu_parity.py -x -q -k "fused or golden" 2>&1 | tail -4 | tee $out/pytest_fused.log
python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_c4.json 2> $out/bench_c4.err; tail -c 2500 $out/bench_c4.json; tail -3 $out/bench_c4.err
for v in default mat_mb2 mat_mb2_noinl mat_noinl; do
  if [ $v = default ]; then unset TUNE_LIB; else export TUNE_LIB=fdtd_b200/_variants/lib_$v.so; fi
  for c in c3 c5slab; do echo "# $v $c"; python scripts/bench_configs.py $c 2>&1 | tail -1 | cut -c1-200; done
done 2>&1 | tee $out/mat_variants.log
unset TUNE_LIB
python - <<'PY' 2>&1 | tee $out/fused_sizes.log
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fdtd_b200 as fd
from bench import build_c4
fd.set_backend("cuda.float32")
for n in (384, 512, 640, 768, 896):
    for fuse in (0, 1):
        g = build_c4(fd, n); g._fuse_eh = fuse
        g.run(4, progress_bar=False); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.run(20, progress_bar=False); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(n, "fused" if fuse else "two half-steps", f"{ms:.3f} ms/step {n**3 / ms / 1e3:.0f} Mcell/s", flush=True)
        del g
PY
