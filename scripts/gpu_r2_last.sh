#!/bin/bash
# the last hardware check of round 2 (development tool, 1 GPU): GPU suite, smoke, the driver's default bench command (timed)
set -u
out=gpurun_out/r2_last
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $out/smoke.log
t0=$(date +%s)
python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench.py (default flags) took $(( $(date +%s) - t0 )) s" | tee $out/bench_time.txt
cut -c1-400 $out/bench_default.json
