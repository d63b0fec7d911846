#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log
timeout 600 python bench.py --rows 1250000 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --verify-queries 2 > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; echo "rc=$?" >> gpurun_out/r2_bench10.err
