#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest16.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err; echo "rc=$?" >> gpurun_out/r2_bench16.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench16_ref.json 2> gpurun_out/r2_bench16_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke16.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke16.log
