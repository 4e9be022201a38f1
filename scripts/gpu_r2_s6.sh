#!/bin/bash
# Round-2 GPU session 6 (development tool, 1 GPU): psi prefetch in the pipelined fused kernel, x-chunk length.
set -u
out=gpurun_out/r2_s6
mkdir -p $out
for lib in fdtd_b200/_variants/lib_pipe_*.so; do
  for xc in 32 64 128; do
    echo "# $lib x_chunk=$xc"
    X_CHUNK=$xc TUNE_LIB=$lib FDTD_B200_FUSE_EH=3 timeout 120 python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
  done
done 2>&1 | tee $out/pipe_psi.jsonl
