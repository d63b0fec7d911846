#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -q -k "int8" > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest20.log
timeout 900 python scripts/bench_configs.py i8_10Mx768_B16_tensor i8_10Mx768_B64_tensor i8_10Mx768_B256_tensor i8_10Mx768_B1024_tensor i8_10Mx768_B4096_tensor > gpurun_out/r2_i8_gang.jsonl 2> gpurun_out/r2_i8_gang.err; echo "rc=$?" >> gpurun_out/r2_i8_gang.err
