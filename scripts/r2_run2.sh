#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
timeout 300 python bench.py --rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench2_1p25M.json 2> gpurun_out/r2_bench2_1p25M.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches2_1p25M.csv \
  python bench.py --rows 1250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --verify-queries 0 > gpurun_out/r2_ncu_l2.log 2>&1
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
