#!/bin/bash
# Round-2 GPU session 10 (development tool, 1 GPU): the fused kernel with TMA staging -- correctness, timing against
# the cp.async build, DRAM bytes.
set -u
out=gpurun_out/r2_s10
mkdir -p $out
python -m pytest tests/test_gpu_parity.py -x -q -k "fused" 2>&1 | tail -4 | tee $out/pytest_fused.log
echo "# default build (TMA)"; python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
for lib in fdtd_b200/_variants/lib_pipe_*.so; do
  echo "# $lib"; TUNE_LIB=$lib python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
echo "# two half-steps"; FDTD_B200_FUSE_EH=0 python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fused_eh_pipe -s 6 -c 1 --csv --log-file $out/tma_metrics.csv python scripts/bench_configs.py c4 > /dev/null 2>&1
grep fused $out/tma_metrics.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
