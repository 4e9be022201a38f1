# multi-GPU confirmation: sharded parity tests + strong-scaling bench lines (+ config 5 on the full box)
N=${1:-2}
STEPS=${2:-40}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$N.txt
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_sharded_$N.log
timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $STEPS --warmup 3 2> gpurun_out/bench_n$n.err | tee gpurun_out/bench_n$n.json
    grep -i "warn\|Traceback" gpurun_out/bench_n$n.err | head -3
  fi
done
if [ $N -ge 8 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 scripts/bench_configs.py c5 2> gpurun_out/c5_n8.err | tee gpurun_out/c5_n8.jsonl
  grep -i "warn\|Traceback\|Error" gpurun_out/c5_n8.err | head -5
fi
true
