#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -q -k "fp32" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
timeout 900 python scripts/bench_configs.py f32_5Mx768_B1_tensor f32_5Mx768_B8_tensor f32_5Mx768_B16_tensor f32_5Mx768_B16_scan f32_5Mx768_B1_scan > gpurun_out/r2_f32_configs.jsonl 2> gpurun_out/r2_f32_configs.err
