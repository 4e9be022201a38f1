#!/bin/bash
# Round-2 GPU session 2 (development tool): ncu --set full of the shipping half-step kernels on configs 2/3/4 and
# of the pipelined fused E+H kernel (7 rows x 31 lanes, 2 blocks per SM) at 1024^3.
set -u
out=gpurun_out/r2_s2
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:halfstep_kernel -s 10 -c 2 -o $out/c4_halfstep python scripts/bench_configs.py c4 > $out/c4.log 2>&1
$NCU -k regex:halfstep_kernel -s 10 -c 2 -o $out/c3_halfstep python scripts/bench_configs.py c3 > $out/c3.log 2>&1
$NCU -k regex:halfstep_kernel -s 10 -c 2 -o $out/c2_halfstep python scripts/bench_configs.py c2 > $out/c2.log 2>&1
TUNE_LIB=fdtd_b200/_variants/lib_pipe_r7l31_mb2.so FDTD_B200_FUSE_EH=3 $NCU -k regex:fused_eh_pipe -s 6 -c 1 \
  -o $out/c4_pipe python scripts/bench_configs.py c4 > $out/c4_pipe.log 2>&1
tail -2 $out/*.log
ls -la $out
