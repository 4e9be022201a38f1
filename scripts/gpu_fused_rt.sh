#!/bin/bash
# Time the register-tiled fused E+H kernel variants at 1024^3 f32 on one B200 (development tool).
#   here (no GPU):   python scripts/tune.py build rt_r2_w4_mb3,rt_r2_w2_mb6,rt_r2_w8_mb1,rt_r3_w4_mb2,rt_r4_w4_mb2,rt_r4_w2_mb4,rt_r1_w4_mb4
#   on the GPU box:  bash scripts/gpu_fused_rt.sh > gpurun_out/fused_rt.jsonl
# (fdtd_b200/_variants/ is git-ignored but travels with the gpurun snapshot; delete it afterwards)
for lib in fdtd_b200/_variants/lib_rt_*.so; do
  echo "# $lib"
  TUNE_LIB=$lib FDTD_B200_FUSE_EH=2 timeout 60 python scripts/bench_configs.py c4 2>&1 | tail -1
done
for lib in fdtd_b200/_variants/lib_pipe_*.so; do
  echo "# $lib"
  TUNE_LIB=$lib FDTD_B200_FUSE_EH=3 timeout 60 python scripts/bench_configs.py c4 2>&1 | tail -1
done
for lib in fdtd_b200/_variants/lib_fz*.so; do
  [ -e "$lib" ] || continue
  echo "# $lib"
  TUNE_LIB=$lib FDTD_B200_FUSE_EH=1 timeout 60 python scripts/bench_configs.py c4 2>&1 | tail -1
done
echo "# two half-steps"
FDTD_B200_FUSE_EH=0 timeout 60 python scripts/bench_configs.py c4 2>&1 | tail -1
echo "# shared-memory fused kernel"
FDTD_B200_FUSE_EH=1 timeout 60 python scripts/bench_configs.py c4 2>&1 | tail -1
echo "# pipelined fused kernel (default build)"
FDTD_B200_FUSE_EH=3 timeout 60 python scripts/bench_configs.py c4 2>&1 | tail -1
