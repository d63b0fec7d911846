#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -q -k "fp32" > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15.log
for v in 1 0; do
B200R_F32_MID_TMEM=$v timeout 600 python scripts/bench_configs.py f32_5Mx768_B16_tensor f32_5Mx768_B1_tensor >> gpurun_out/r2_f32_tmem.jsonl 2>> gpurun_out/r2_f32_tmem.err
done
