#!/bin/bash
# Round-2 4-GPU session (development tool; gpurun --gpus 4): interior ranks (two neighbours) of the split fused sharded
# step -- sharded == single on 4 ranks, bench c4 at N=4, thin slabs (128 planes per rank).
set -u
out=gpurun_out/r2_n4c
mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_sharded.py -x -q -k "fused or c4small-float32-p2p" 2>&1 | tail -4 | tee $out/pytest_sharded_4ranks.log
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
{
FDTD_B200_FUSE_SPLIT=1 $TR scripts/slab_bench.py 512 40 2>/dev/null | tail -1
FDTD_B200_FUSE_SPLIT=0 $TR scripts/slab_bench.py 512 40 2>/dev/null | tail -1
} | tee $out/slab_bench.log
$TR bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $out/bench_c4_n4.json
cut -c1-330 $out/bench_c4_n4.json
