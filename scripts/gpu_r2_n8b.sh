#!/bin/bash
# Round-2 8-GPU session b (development tool; gpurun --gpus 8): c4 strong-scaling curve with the TMA-staged fused kernel.
set -u
out=gpurun_out/r2_n8b
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -k "fused or slices" 2>&1 | tail -3 | tee $out/pytest_fused_4ranks.log
for N in 8 4 2 1; do
  if [ $N = 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 600 $RUN bench.py --gpus $N --config c4 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_c4_n$N.json 2> $out/bench_c4_n$N.err
  python - <<PY
import json
try:
    l = json.loads(open("$out/bench_c4_n$N.json").read().strip().splitlines()[-1])
    p = l.get("parity") or {}
    print("c4 N=$N", round(l["value"]), "Mcell/s", round(l["ms_per_step"], 4), "ms e2e", round(l["e2e"]["value"]), "step frac", round(l["hbm_roofline_frac_whole_step"], 4),
          "parity", p.get("sharded_equals_single"), (p.get("fused_steps") or {}).get("active"), "launches", l["gpu_launches"], flush=True)
except Exception as e:
    print("c4 N=$N ERR", e, open("$out/bench_c4_n$N.err").read()[-800:])
PY
done 2>&1 | tee $out/curve_c4.txt
FDTD_B200_FUSE_EH=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config c4 --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_c4_n8_twopass.json 2> $out/bench_c4_n8_twopass.err
python -c "
import json; l=json.loads(open('$out/bench_c4_n8_twopass.json').read().strip().splitlines()[-1]); print('c4 N=8 two half-steps', round(l['value']), round(l['ms_per_step'],4), 'e2e', round(l['e2e']['value']))" | tee -a $out/curve_c4.txt
