#!/bin/bash
# Round-2 GPU session 13 (development tool, 1 GPU): the fused kernel with the z+1 neighbour by shuffle and the z slab's
# psi staged by bulk copies -- correctness on hardware, timing of each change alone, source-level ncu capture.
set -u
out=gpurun_out/r2_s13
mkdir -p $out /tmp/rep
python -m pytest tests/test_gpu_parity.py -x -q -k "fused" 2>&1 | tail -4 | tee $out/pytest_fused.log
{
echo "# default build (shuffle + staged psi)"; python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
for lib in fdtd_b200/_variants/lib_pipe_v2_*.so; do
  echo "# $lib"; TUNE_LIB=$lib python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
echo "# default build, x_chunk=64"; X_CHUNK=64 python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
echo "# default build, x_chunk=32"; X_CHUNK=32 python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
} | tee $out/variants.log
ncu --set full --clock-control none --import-source on -k regex:fused_eh_pipe -s 6 -c 1 -o /tmp/rep/c4_fused_v2 python scripts/bench_configs.py c4 > $out/c4_fused_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/rep/c4_fused_v2.ncu-rep > $out/ncu_summary_fused_v2.txt 2>&1
cp /tmp/rep/c4_fused_v2.ncu-rep $out/
tail -25 $out/ncu_summary_fused_v2.txt
