mkdir -p gpurun_out
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 5 "$@" 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('N=%d grid=%s value=%.0f ms/step=%.4f e2e=%.0f kfrac=%.3f kernel_ms=%.4f per_rank=%s' % (d['n_gpus'], d['config']['grid'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d.get('per_rank_ms_per_step')))
"; }
echo "== N=1 128x1024x1024"; python bench.py --steps 40 --warmup 5 --shape 128,1024,1024 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('value=%.0f ms/step=%.4f kfrac=%.3f kernel_ms=%.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch']))"
echo "== N=2 256x1024x1024"; run2 --shape 256,1024,1024
echo "== N=2 256x1024x1024 skip halo"; FDTD_B200_SKIP_HALO=1 run2 --shape 256,1024,1024
echo "== N=2 1024^3"; run2
