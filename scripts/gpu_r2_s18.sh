#!/bin/bash
# Round-2 GPU session 18 (development tool, 1 GPU): evict-first stores of the fused kernel's results; x-chunk lengths
# around the chosen one; DRAM bytes of the variant.
set -u
out=gpurun_out/r2_s18
mkdir -p $out
{
echo "# default build"; python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
for lib in fdtd_b200/_variants/lib_pipe_v6_*.so; do
  echo "# $lib"; TUNE_LIB=$lib python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
for xc in 40 56; do
  echo "# default build, x_chunk=$xc"; X_CHUNK=$xc python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
} | tee $out/variants.log
for lib in fdtd_b200/_variants/lib_pipe_v6_cs.so; do
TUNE_LIB=$lib ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_eh_pipe -s 6 -c 1 --csv --log-file $out/cs_metrics.csv python scripts/bench_configs.py c4 > /dev/null 2>&1
grep fused $out/cs_metrics.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
