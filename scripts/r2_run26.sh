#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -q -k "fp32 or f32" > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest26.log
# fp32 tensor kernel: hi piece in tensor memory (default) against hi in shared memory, same box
for h in 1 0; do
  B200R_F32_HI_TMEM=$h timeout 600 python scripts/bench_configs.py f32_5Mx768_B1_tensor f32_5Mx768_B16_tensor > gpurun_out/r2_f32_hitmem$h.jsonl 2> gpurun_out/r2_f32_hitmem$h.err
done
