#!/bin/bash
# Round-2 GPU session 15 (development tool, 1 GPU): fused kernel after the clean-up (block barrier kept, tables in shared
# memory) -- correctness, TMA L2 prefetch variants, racecheck, ncu capture of the shipping build.
set -u
out=gpurun_out/r2_s15
mkdir -p $out /tmp/rep
python -m pytest tests/test_gpu_parity.py -x -q -k "fused" 2>&1 | tail -4 | tee $out/pytest_fused.log
{
echo "# default build"; python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
for lib in fdtd_b200/_variants/lib_pipe_v4_*.so; do
  echo "# $lib"; TUNE_LIB=$lib python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done
echo "# default build again"; python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
} | tee $out/variants.log
for tool in racecheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_scenes.py only-fused > $out/$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|error" $out/$tool.log | tail -5
done
ncu --set full --clock-control none --import-source on -k regex:fused_eh_pipe -s 6 -c 1 -o /tmp/rep/c4_fused_v4 python scripts/bench_configs.py c4 > $out/c4_fused_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/rep/c4_fused_v4.ncu-rep > $out/ncu_summary_fused_v4.txt 2>&1
cp /tmp/rep/c4_fused_v4.ncu-rep $out/
tail -8 $out/ncu_summary_fused_v4.txt
