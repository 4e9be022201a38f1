mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_p2p.log
for mode in p2p nccl; do
  echo "== halo=$mode 1024^3 N=2"
  FDTD_B200_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 40 --warmup 5 2>gpurun_out/p2p_$mode.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('N=%d grid=%s value=%.0f ms/step=%.4f e2e=%.0f per_rank=%s launches=%d' % (d['n_gpus'], d['config']['grid'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('per_rank_ms_per_step'), d['gpu_launches']))
"
  echo "== halo=$mode 256x1024x1024 N=2"
  FDTD_B200_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 40 --warmup 5 --shape 256,1024,1024 2>>gpurun_out/p2p_$mode.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('N=%d grid=%s value=%.0f ms/step=%.4f e2e=%.0f per_rank=%s launches=%d' % (d['n_gpus'], d['config']['grid'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('per_rank_ms_per_step'), d['gpu_launches']))
"
  grep -i "warn\|error\|Traceback" gpurun_out/p2p_$mode.err | head -5
done
