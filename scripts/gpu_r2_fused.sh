#!/bin/bash
# Round-2 first GPU session for the temporally fused E+H kernels (development tool; one B200).
#   here (no GPU):  python scripts/tune.py build rt_r2_w4_mb3,rt_r2_w2_mb6,rt_r3_w4_mb2,rt_r4_w2_mb4,pipe_r4l32_mb3,pipe_r4l32_mb2,pipe_r8l32_mb1,pipe_r8l16_mb2,pipe_r2l32_mb4,pipe_r4l16_mb4,pipe_r4l31_mb3,pipe_r4l31_mb2,pipe_r7l31_mb2,fz_r4l31_mb4
#   on the box:     gpurun --timeout 900 -- 'bash scripts/gpu_r2_fused.sh'
# Writes everything under gpurun_out/r2_fused/.
set -u
out=gpurun_out/r2_fused
mkdir -p $out
# 1. correctness of every variant on hardware, and ms per step at 512^3
python scripts/gpu_fused_check.py 1 2 3 --time 2>&1 | tee $out/check.log
# 2. 1024^3 f32: two half-steps, the three variants of the default build, then the build variants
for f in 0 1 2 3; do
  echo "# FUSE_EH=$f (default build)"
  FDTD_B200_FUSE_EH=$f timeout 90 python scripts/bench_configs.py c4 2>&1 | tail -1
done | tee $out/c4_default.jsonl
if ls fdtd_b200/_variants/lib_*.so >/dev/null 2>&1; then bash scripts/gpu_fused_rt.sh 2>&1 | tee $out/c4_variants.jsonl; fi
# 3. launch list of a few fused steps (shares of the step), then one full capture of the pipelined kernel:
#    achieved occupancy (is it 3 blocks per SM?), dram bytes, stall reasons per source line
FDTD_B200_FUSE_EH=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file $out/launches_pipe.csv python scripts/bench_configs.py c2 > $out/launches_pipe.log 2>&1
FDTD_B200_FUSE_EH=3 ncu --set full --clock-control none --import-source on -k regex:fused_eh_pipe -s 4 -c 2 \
  -o $out/pipe python scripts/gpu_fused_check.py 3 --time > $out/ncu_pipe.log 2>&1
FDTD_B200_FUSE_EH=1 ncu --set full --clock-control none --import-source on -k regex:fused_eh_kernel -s 4 -c 2 \
  -o $out/smem python scripts/gpu_fused_check.py 1 --time > $out/ncu_smem.log 2>&1
ls -la $out
