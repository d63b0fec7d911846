#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -q -k "fp32 or f32" > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest27.log
timeout 600 python scripts/bench_configs.py f32_5Mx768_B1_tensor f32_5Mx768_B16_tensor f32_5Mx768_B64_tensor > gpurun_out/r2_f32_kparts.jsonl 2> gpurun_out/r2_f32_kparts.err
B200R_F32_HI_TMEM=1 timeout 600 python scripts/bench_configs.py f32_5Mx768_B16_tensor > gpurun_out/r2_f32_kparts_hitmem.jsonl 2> gpurun_out/r2_f32_kparts_hitmem.err
