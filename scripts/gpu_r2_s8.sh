#!/bin/bash
# Round-2 GPU session 8 (development tool, 1 GPU): ncu captures of the SHIPPING kernels for profiles/ (full sets +
# the launch list of the bench command), after prefetch cap / plane-run split / MAT occupancy / fused default.
set -u
out=gpurun_out/r2_s8
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:fused_eh_pipe -s 6 -c 1 -o $out/c4_fused python scripts/bench_configs.py c4 > $out/c4_fused.log 2>&1
FDTD_B200_FUSE_EH=0 $NCU -k regex:halfstep_kernel -s 10 -c 2 -o $out/c4_halfstep python scripts/bench_configs.py c4 > $out/c4_halfstep.log 2>&1
$NCU -k regex:halfstep_kernel -s 24 -c 8 -o $out/c3_halfstep python scripts/bench_configs.py c3 > $out/c3.log 2>&1
$NCU -k regex:halfstep_kernel -s 10 -c 2 -o $out/c2_halfstep python scripts/bench_configs.py c2 > $out/c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_c4.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_under_ncu.log 2>&1
tail -n 2 $out/*.log | cut -c1-300
ls -la $out
