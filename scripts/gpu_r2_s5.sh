#!/bin/bash
# Round-2 GPU session 5 (development tool, 1 GPU): pipelined fused E+H kernel variants at 1024^3 f32.
set -u
out=gpurun_out/r2_s5
mkdir -p $out
echo "# two half-steps (default build, prefetch capped)"; python scripts/bench_configs.py c4 2>&1 | tail -1
for lib in fdtd_b200/_variants/lib_pipe_*.so; do
  echo "# $lib"
  TUNE_LIB=$lib FDTD_B200_FUSE_EH=3 timeout 120 python scripts/bench_configs.py c4 2>&1 | tail -1
done 2>&1 | tee $out/pipe_variants.jsonl
