#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -q -k "fp32 or f32 or float32 or scan_bulk or bulk" > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest23.log
timeout 900 python scripts/bench_configs.py f32_5Mx768_B16_tensor f32_5Mx768_B64_tensor f32_5Mx768_B256_tensor > gpurun_out/r2_f32_gang.jsonl 2> gpurun_out/r2_f32_gang.err; echo "rc=$?" >> gpurun_out/r2_f32_gang.err
# batch 1 / batch 16 on the N=8 shard size: per-launch-site timeline of the small-batch step
for b in 1 16; do
  timeout 600 python bench.py --rows 1250000 --batch $b --steps 200 --warmup 20 --no-cpu-baseline --no-legs --no-e2e --verify-queries 1 > gpurun_out/r2_bench23_shard_b$b.json 2> gpurun_out/r2_bench23_shard_b$b.err
done
