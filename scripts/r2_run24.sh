#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest24.log
# fp32 tensor kernel: 8 mid buffers in tensor memory (default) against the 3 it had
for m in 8 3 5; do
  B200R_F32_MID_BUFS=$m timeout 600 python scripts/bench_configs.py f32_5Mx768_B1_tensor f32_5Mx768_B16_tensor > gpurun_out/r2_f32_mids$m.jsonl 2> gpurun_out/r2_f32_mids$m.err
done
for b in 1 16; do
  timeout 600 python bench.py --rows 1250000 --batch $b --steps 200 --warmup 20 --no-cpu-baseline --no-legs --no-e2e --verify-queries 1 > gpurun_out/r2_bench24_shard_b$b.json 2> gpurun_out/r2_bench24_shard_b$b.err
done
