mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_c.log
timeout 900 python scripts/bench_configs.py 2>&1 | grep -v Warning | tee gpurun_out/configs_r1.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_fused.json 2> gpurun_out/bench_r1_fused.err; cat gpurun_out/bench_r1_fused.json; tail -3 gpurun_out/bench_r1_fused.err
