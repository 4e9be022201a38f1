#!/bin/bash
# Round-2 GPU session 12 (development tool): x-chunk length of the TMA-staged fused kernel at 1024^3 f32.
mkdir -p gpurun_out/r2_s12
for xc in 8 16 24 32 48 64 128; do
  echo "# x_chunk=$xc"; X_CHUNK=$xc python scripts/bench_configs.py c4 2>&1 | tail -1 | cut -c1-140
done | tee gpurun_out/r2_s12/xchunk_tma.log
