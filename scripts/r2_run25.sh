#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest25.log
for b in 1 16; do
  timeout 600 python bench.py --rows 1250000 --batch $b --steps 200 --warmup 20 --no-cpu-baseline --no-legs --no-e2e --verify-queries 1 > gpurun_out/r2_bench25_shard_b$b.json 2> gpurun_out/r2_bench25_shard_b$b.err
done
timeout 600 python bench.py --batch 1 --steps 50 --warmup 10 --no-cpu-baseline --no-legs --no-e2e --verify-queries 1 > gpurun_out/r2_bench25_10M_b1.json 2> gpurun_out/r2_bench25_10M_b1.err
