#!/bin/bash
# Round-2 multi-GPU session (development tool; gpurun --gpus N): sharded == single on hardware (peer-to-peer and NCCL
# halo), the C-side sharded loop, bench lines with the parity field for c4 and c5.
set -u
N=${1:-2}
out=gpurun_out/r2_n$N
mkdir -p $out
nvidia-smi -L | tee $out/gpus.txt
python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -6 | tee $out/pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for c in c4 c5; do
  $TR bench.py --gpus $N --config $c --steps 20 --warmup 3 > $out/bench_$c.json 2> $out/bench_$c.err
  tail -c 2200 $out/bench_$c.json; tail -3 $out/bench_$c.err
done
$TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
tail -c 700 $out/bench_ref.json
