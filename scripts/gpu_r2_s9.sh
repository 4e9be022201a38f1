#!/bin/bash
# Round-2 GPU session 9 (development tool, 1 GPU): concurrent plane-run launches (c3), GPU suite, launch list of c3.
set -u
out=gpurun_out/r2_s9
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest_gpu.log
for c in c3 c5slab c2; do python scripts/bench_configs.py $c 2>&1 | tail -1 | cut -c1-200; done | tee $out/configs.jsonl
python bench.py --config c3 --steps 20 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 | cut -c1-900 | tee $out/bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 -s 60 --csv --log-file $out/launches_c3.csv python scripts/bench_configs.py c3 > /dev/null 2>&1
grep halfstep $out/launches_c3.csv | awk -F'","' '{print $5, $NF}' | head -16
