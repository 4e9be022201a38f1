#!/bin/bash
# final check of the tree on hardware (development tool; gpurun --gpus 2): GPU suite incl. the multi-GPU cases, smoke,
# default bench at N=1 and N=2
set -u
out=gpurun_out/r2_check
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest_gpu_2gpus.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 | cut -c1-330 | tee $out/bench_n1.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-330 | tee $out/bench_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
