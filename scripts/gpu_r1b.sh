mkdir -p gpurun_out
echo "== slab 128x1024x1024" | tee gpurun_out/tune4.log
TUNE_SIZE=128x1024x1024 TUNE_CHUNKS=4,8,16,32,64 python scripts/tune.py run mb3_pf1 2>&1 | grep -v Warning | tee -a gpurun_out/tune4.log
echo "== slab 256x1024x1024" | tee -a gpurun_out/tune4.log
TUNE_SIZE=256x1024x1024 TUNE_CHUNKS=8,16,32 python scripts/tune.py run mb3_pf1 2>&1 | grep -v Warning | tee -a gpurun_out/tune4.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_tuned.json 2> gpurun_out/bench_r1_tuned.err
cat gpurun_out/bench_r1_tuned.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1_tuned.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:halfstep -s 6 -c 2 -f -o gpurun_out/prof_r1_tuned python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_ncu2.log 2>&1
tail -2 gpurun_out/prof_ncu2.log
