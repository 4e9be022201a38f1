#!/bin/bash
# Round-2 GPU session 3 (development tool): GPU test-suite after the ABI-15 / float32x changes, float32x speed,
# sharded tests and C-side sharded loop on 2 GPUs.
set -u
out=gpurun_out/r2_s3
mkdir -p $out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 | tee $out/pytest_gpu.log
python - <<'PY' 2>&1 | tee $out/f32x_speed.log
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fdtd_b200 as fd
from bench import build_c4
for mode, n in (("float32", 512), ("float32x", 512), ("float32", 1024), ("float32x", 1024)):
    fd.set_backend("cuda." + mode)
    g = build_c4(fd, n)
    g.run(3, progress_bar=False); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.run(20, progress_bar=False); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(mode, n, f"{ms:.3f} ms/step, {n**3 / ms / 1e3:.0f} Mcell/s", flush=True)
    del g
PY
for c in c4 c3 c2 c1; do
  python bench.py --config $c --steps 20 --warmup 3 > $out/bench_$c.json 2> $out/bench_$c.err
  tail -c 1500 $out/bench_$c.json; tail -3 $out/bench_$c.err
done
python bench.py --config c4 --mode float32x --steps 10 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_c4_f32x.json 2> $out/bench_c4_f32x.err
tail -c 600 $out/bench_c4_f32x.json
