mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_d.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_d.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; cat gpurun_out/bench_r1_final.json; tail -3 gpurun_out/bench_r1_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu3.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:halfstep -s 6 -c 2 -f -o gpurun_out/prof_r1_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_ncu3.log 2>&1
tail -2 gpurun_out/prof_ncu3.log
timeout 600 python scripts/bench_configs.py c1 c2 c3 c5slab 2>&1 | grep -v Warning | tee gpurun_out/configs_r1c.jsonl
