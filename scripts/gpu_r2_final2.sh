#!/bin/bash
# Round-2 last 1-GPU session (development tool): full GPU suite + smoke, bench lines of configs 1-4 with all baselines,
# ncu --set full of the TMA-staged fused kernel (summarised on the box), launch list of the bench command.
set -u
out=gpurun_out/r2_final2
mkdir -p $out /tmp/rep
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -12 | tee $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $out/smoke.log
python bench.py --config c4 --steps 20 --warmup 5 > $out/bench_c4_n1.json 2> $out/bench_c4.err
python bench.py --config c3 --steps 50 --warmup 5 > $out/bench_c3_n1.json 2> $out/bench_c3.err
python bench.py --config c2 --steps 200 --warmup 40 > $out/bench_c2_n1.json 2> $out/bench_c2.err
python bench.py --config c1 --steps 1000 --warmup 100 > $out/bench_c1_n1.json 2> $out/bench_c1.err
python bench.py --config c4 --mode float32x --steps 10 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_c4_f32x_n1.json 2> $out/bench_c4_f32x.err
for c in c4 c3 c2 c1 c4_f32x; do python - <<PY
import json
try:
    l = json.loads(open("$out/bench_${c}_n1.json").read().strip().splitlines()[-1])
    print("$c", round(l["value"]), "Mcell/s", round(l["ms_per_step"], 4), "ms; e2e", round(l["e2e"]["value"]), "; step frac", round(l["hbm_roofline_frac_whole_step"], 4), "; kernel", l["roofline"]["kernel"][:22], round(l["roofline"]["frac"], 4), "; cpu", (l.get("cpu_baseline") or {}).get("value"), "eager", (l.get("gpu_eager_baseline") or {}).get("value"))
except Exception as e:
    print("$c ERR", e, open("$out/bench_${c%_f32x}.err").read()[-600:])
PY
done | tee $out/summary.txt
ncu --set full --clock-control none --import-source on -k regex:fused_eh_pipe -s 6 -c 1 -o /tmp/rep/c4_fused_final python scripts/bench_configs.py c4 > $out/ncu_fused.log 2>&1
python scripts/ncu_summary.py /tmp/rep/c4_fused_final.ncu-rep > $out/ncu_summary_fused_final.txt 2>&1
ncu -i /tmp/rep/c4_fused_final.ncu-rep --page raw --csv > $out/c4_fused_final_raw.csv 2>/dev/null
cuobjdump -sass fdtd_b200/libfdtd_b200.so 2>/dev/null | grep -E "UTMALDG|UBLKCP|SYNCS|SHFL" | awk '{print $2}' | sort | uniq -c > $out/sass_tma.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_c4.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > $out/bench_under_ncu.log 2>&1
cat $out/ncu_summary_fused_final.txt | head -30; du -sh $out
