mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|float32 vs float64|worst rel" | tail -30 | tee gpurun_out/pytest_gpu_verify.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_verify.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; cat gpurun_out/bench_verify.json; tail -3 gpurun_out/bench_verify.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_arm.json
timeout 600 python scripts/bench_configs.py c1 c2 c3 c3b c5slab 2>&1 | grep -v Warning | tee gpurun_out/configs_verify.jsonl
