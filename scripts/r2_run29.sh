#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest29.log
timeout 600 python scripts/bench_configs.py f32_5Mx768_B1_tensor f32_5Mx768_B16_tensor > gpurun_out/r2_f32_final.jsonl 2> gpurun_out/r2_f32_final.err
# ncu of the two selects of a batch-1 search on the N=8 shard size (one wide CTA each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_lists -s 4 -c 2 -o gpurun_out/r2_select_b1 -f \
  python bench.py --rows 1250000 --batch 1 --steps 2 --warmup 1 --no-cpu-baseline --no-legs --no-e2e --verify-queries 0 > gpurun_out/r2_ncu_select_b1.log 2>&1
