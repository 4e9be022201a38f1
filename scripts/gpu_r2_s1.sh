#!/bin/bash
# Round-2 GPU session 1 (development tool): state of the tree on hardware + the fused E+H variants.
set -u
out=gpurun_out/r2_s1
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $out/gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/pytest_gpu.log
python scripts/gpu_fused_check.py 1 3 --time 2>&1 | tee $out/check.log
for f in 0 1 3; do
  echo "# FUSE_EH=$f (default build)"
  FDTD_B200_FUSE_EH=$f timeout 120 python scripts/bench_configs.py c4 2>&1 | tail -1
done | tee $out/c4_default.jsonl
for lib in fdtd_b200/_variants/lib_pipe_*.so; do
  echo "# $lib"
  TUNE_LIB=$lib FDTD_B200_FUSE_EH=3 timeout 120 python scripts/bench_configs.py c4 2>&1 | tail -1
done | tee $out/c4_variants.jsonl
FDTD_B200_FUSE_EH=3 ncu --set full --clock-control none --import-source on -k regex:fused_eh_pipe -s 4 -c 1 \
  -o $out/pipe python scripts/gpu_fused_check.py 3 --time > $out/ncu_pipe.log 2>&1
ls -la $out
