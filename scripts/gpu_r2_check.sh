#!/bin/bash
# final check of the tree on hardware (development tool; gpurun --gpus 2): GPU suite incl. the multi-GPU cases, smoke,
# default bench at N=1 and N=2, the reference arm, config 5 on two GPUs
set -u
out=gpurun_out/r2_check
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest_gpu_2gpus.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/smoke.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline 2>/dev/null | tail -1 > $out/bench_n1.json; cut -c1-330 $out/bench_n1.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 > $out/bench_n2.json; cut -c1-330 $out/bench_n2.json
$TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
$TR bench.py --config c5 --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 > $out/bench_c5_n2.json; cut -c1-330 $out/bench_c5_n2.json
