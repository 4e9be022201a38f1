# multi-GPU confirmation: NCCL sharded parity tests + strong-scaling bench lines
N=${1:-2}
STEPS=${2:-40}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$N.txt
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_sharded_$N.log
timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $STEPS --warmup 3 2> gpurun_out/bench_n$n.err | tee gpurun_out/bench_n$n.json
    tail -3 gpurun_out/bench_n$n.err | grep -v "^\*\|OMP\|^$"
  fi
done
