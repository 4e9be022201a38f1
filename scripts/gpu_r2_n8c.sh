#!/bin/bash
# Round-2 8-GPU session (development tool; gpurun --gpus 8): bench c4 at N=8 with the final kernels.
set -u
out=gpurun_out/r2_n8c
mkdir -p $out
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
$TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline 2>$out/bench_c4_n8.err | tail -1 > $out/bench_c4_n8.json
cut -c1-330 $out/bench_c4_n8.json; tail -3 $out/bench_c4_n8.err
