#!/bin/bash
# NVLink bytes of the fused compute + push kernel (development tool; gpurun --gpus 2, ONE process)
mkdir -p gpurun_out/r2_nvlink
python scripts/nvlink_push_probe.py 2>&1 | tail -2
ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:halfstep_kernel --csv --log-file gpurun_out/r2_nvlink/push_metrics.csv python scripts/nvlink_push_probe.py > gpurun_out/r2_nvlink/probe.log 2>&1
tail -2 gpurun_out/r2_nvlink/probe.log
grep -E "halfstep" gpurun_out/r2_nvlink/push_metrics.csv | awk -F'","' '{print $5, "|", $(NF-2), $(NF-1), $NF}' | tail -24
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl" | head -12
