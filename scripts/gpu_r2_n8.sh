#!/bin/bash
# Round-2 8-GPU session (development tool; gpurun --gpus 8): sharded tests on 4 ranks, the 1/2/4/8 curves of c4
# (strong) and c5 (weak, 256 planes per GPU).
set -u
out=gpurun_out/r2_n8
mkdir -p $out
nvidia-smi -L | wc -l | tee $out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -5 | tee $out/pytest_sharded_4ranks.log
for c in c4 c5; do
  for N in 8 4 2 1; do
    if [ $N = 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
    timeout 600 $RUN bench.py --gpus $N --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_${c}_n$N.json 2> $out/bench_${c}_n$N.err
    python - <<PY
import json
try:
    l = json.loads(open("$out/bench_${c}_n$N.json").read().strip().splitlines()[-1])
    p = l.get("parity") or {}
    print("$c N=$N", round(l["value"]), "Mcell/s", round(l["ms_per_step"], 4), "ms e2e", round(l["e2e"]["value"]), "step frac", round(l["hbm_roofline_frac_whole_step"], 4),
          "kernel", l["roofline"]["kernel"][:24], round(l["roofline"]["frac"], 4), "parity", p.get("sharded_equals_single"), (p.get("fused_steps") or {}).get("active"), flush=True)
except Exception as e:
    print("$c N=$N ERR", e, open("$out/bench_${c}_n$N.err").read()[-800:])
PY
  done
done 2>&1 | tee $out/curves.txt
FDTD_B200_FUSE_EH=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config c4 --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_c4_n8_twopass.json 2> $out/bench_c4_n8_twopass.err
python -c "
import json; l=json.loads(open('$out/bench_c4_n8_twopass.json').read().strip().splitlines()[-1]); print('c4 N=8 two half-steps', round(l['value']), round(l['ms_per_step'],4))" | tee -a $out/curves.txt
