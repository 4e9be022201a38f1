#!/bin/bash
# Round-2 GPU session 8 (development tool, 1 GPU): ncu captures of the SHIPPING kernels for profiles/ (full sets,
# summarised on the box: gpurun_out/ may carry 64 MiB back) and the launch list of the bench command.
set -u
out=gpurun_out/r2_s8
mkdir -p $out /tmp/rep
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:fused_eh_pipe -s 6 -c 1 -o /tmp/rep/c4_fused python scripts/bench_configs.py c4 > $out/c4_fused.log 2>&1
FDTD_B200_FUSE_EH=0 $NCU -k regex:halfstep_kernel -s 10 -c 2 -o /tmp/rep/c4_halfstep python scripts/bench_configs.py c4 > $out/c4_halfstep.log 2>&1
$NCU -k regex:halfstep_kernel -s 24 -c 8 -o /tmp/rep/c3_halfstep python scripts/bench_configs.py c3 > $out/c3.log 2>&1
$NCU -k regex:halfstep_kernel -s 10 -c 2 -o /tmp/rep/c2_halfstep python scripts/bench_configs.py c2 > $out/c2.log 2>&1
python scripts/ncu_summary.py /tmp/rep/c4_fused.ncu-rep /tmp/rep/c4_halfstep.ncu-rep /tmp/rep/c3_halfstep.ncu-rep /tmp/rep/c2_halfstep.ncu-rep > $out/ncu_summary.txt 2>&1
for r in c4_fused c4_halfstep c3_halfstep c2_halfstep; do
  ncu -i /tmp/rep/$r.ncu-rep --page raw --csv > $out/${r}_raw.csv 2>/dev/null
done
cp /tmp/rep/c4_fused.ncu-rep $out/
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_c4.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_under_ncu.log 2>&1
FDTD_B200_FUSE_EH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_c4_twopass.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_under_ncu2.log 2>&1
for lib in fdtd_b200/_variants/lib_pipe_psi*.so; do
  echo "# $lib"; TUNE_LIB=$lib python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done | tee $out/psi_variants.log
du -sh $out
