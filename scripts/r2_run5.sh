#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "rc=$?" >> gpurun_out/r2_bench5.err
