#!/bin/bash
# Round-2 2-GPU session (development tool; gpurun --gpus 2): the fused sharded step with its boundary chunks on the
# high-priority side stream -- correctness (sharded == single), thin-slab timing with / without the split, bench N=2.
set -u
out=gpurun_out/r2_n2c
mkdir -p $out
python -m pytest tests/test_gpu_sharded.py -x -q -k "fused or c4small-float32-p2p or pml3d-float64-p2p" 2>&1 | tail -4 | tee $out/pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
{
FDTD_B200_FUSE_SPLIT=1 $TR scripts/slab_bench.py 256 40 2>/dev/null | tail -1
FDTD_B200_FUSE_SPLIT=0 $TR scripts/slab_bench.py 256 40 2>/dev/null | tail -1
FDTD_B200_FUSE_EH=0 $TR scripts/slab_bench.py 256 40 2>/dev/null | tail -1
FDTD_B200_FUSE_SPLIT=1 $TR scripts/slab_bench.py 512 40 2>/dev/null | tail -1
FDTD_B200_FUSE_SPLIT=0 $TR scripts/slab_bench.py 512 40 2>/dev/null | tail -1
} | tee $out/slab_bench.log
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $out/bench_c4_n2.json
cut -c1-400 $out/bench_c4_n2.json
