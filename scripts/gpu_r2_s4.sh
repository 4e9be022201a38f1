#!/bin/bash
# Round-2 GPU session 4 (development tool, 1 GPU): full GPU suite, per-config bench lines, kernel variants + DRAM bytes.
set -u
out=gpurun_out/r2_s4
mkdir -p $out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -45 | tee $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $out/smoke.log
for c in c3 c2 c1; do
  python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_$c.json 2> $out/bench_$c.err
  python - <<PY
import json
try:
    l = json.loads(open("$out/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", l["value"], l["ms_per_step"], "kernel frac", l["roofline"]["frac"], "step frac", l["hbm_roofline_frac_whole_step"], "e2e", l["e2e"]["value"])
except Exception as e:
    print("$c ERR", e, open("$out/bench_$c.err").read()[-600:])
PY
done
echo "# c2 with folded sources + graphs forced"
FDTD_B200_GRAPHS=1 FDTD_B200_FUSE=1 python scripts/bench_configs.py c2 2>&1 | tail -1 | tee $out/c2_forced.jsonl
echo "# c2 default"
python scripts/bench_configs.py c2 2>&1 | tail -1 | tee -a $out/c2_forced.jsonl
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:halfstep_kernel -s 10 -c 2 --csv"
for v in default pfcap pfcap_pf2 nopf; do
  echo "# variant $v"
  if [ $v = default ]; then unset TUNE_LIB; else export TUNE_LIB=fdtd_b200/_variants/lib_$v.so; fi
  python scripts/bench_configs.py c4 2>&1 | tail -1
  ncu $M --log-file $out/dram_$v.csv python scripts/bench_configs.py c4 > /dev/null 2>&1
  grep -E "halfstep" $out/dram_$v.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -6
done 2>&1 | tee $out/variants.log
unset TUNE_LIB
echo "# x_chunk 64 / 16"
X_CHUNK=64 python scripts/bench_configs.py c4 2>&1 | tail -1 | tee $out/xchunk.log
X_CHUNK=16 python scripts/bench_configs.py c4 2>&1 | tail -1 | tee -a $out/xchunk.log
ls $out
