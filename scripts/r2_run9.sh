#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_comm.py -m gpu -q -k "cuda_graph or virtual_ranks_exchange" > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
timeout 600 python bench.py --rows 1250000 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --verify-queries 0 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "rc=$?" >> gpurun_out/r2_bench9.err
