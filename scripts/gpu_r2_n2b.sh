#!/bin/bash
# Round-2 multi-GPU session b (development tool): temporally fused steps on slabs.
set -u
N=${1:-2}
out=gpurun_out/r2_n${N}b
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -k "fused" 2>&1 | tail -15 | tee $out/pytest_fused_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for f in 2 0; do
  echo "# FDTD_B200_FUSE_EH=$f"
  FDTD_B200_FUSE_EH=$f timeout 600 $TR bench.py --gpus $N --config c4 --steps 20 --warmup 3 > $out/bench_c4_fuse$f.json 2> $out/bench_c4_fuse$f.err
  python - <<PY
import json
try:
    l = json.loads(open("$out/bench_c4_fuse$f.json").read().strip().splitlines()[-1])
    print(l["value"], l["ms_per_step"], "e2e", l["e2e"]["value"], "launches", l["gpu_launches"], l["parity"], l["per_rank_ms_per_step"])
except Exception as e:
    print("ERR", e, open("$out/bench_c4_fuse$f.err").read()[-1500:])
PY
done
