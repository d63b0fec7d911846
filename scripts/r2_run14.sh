#!/bin/bash
mkdir -p gpurun_out
for b in 1 0; do
B200R_F32_BACKOFF=$b timeout 600 python scripts/bench_configs.py f32_5Mx768_B16_tensor >> gpurun_out/r2_f32_backoff.jsonl 2>> gpurun_out/r2_f32_backoff.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_f32 -s 5 -c 1 -o gpurun_out/r2_rows_f32_main -f \
  python scripts/bench_configs.py f32_5Mx768_B16_tensor > gpurun_out/r2_ncu_rows_f32.log 2>&1
