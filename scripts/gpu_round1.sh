mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout 600 python bench.py --steps 10 --warmup 3 --size 512 --no-cpu-baseline > gpurun_out/bench512.json 2>> gpurun_out/bench1.err
timeout 600 python bench.py --steps 10 --warmup 3 --size 256 --dtype float64 --no-cpu-baseline > gpurun_out/bench256f64.json 2>> gpurun_out/bench1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:halfstep -s 6 -c 2 -f -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_ncu.log 2>&1
cat gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
