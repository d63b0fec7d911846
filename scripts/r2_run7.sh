#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
for ratio in 8 4 16; do
B200R_PHASE_RATIO=$ratio timeout 300 python bench.py --rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline --no-legs > gpurun_out/r2_bench7_1p25M_r$ratio.json 2> gpurun_out/r2_bench7_1p25M_r$ratio.err
done
B200R_SELECT_WARP=0 timeout 300 python bench.py --rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline --no-legs > gpurun_out/r2_bench7_1p25M_nowarp.json 2> gpurun_out/r2_bench7_1p25M_nowarp.err
for ratio in 8 4; do
B200R_PHASE_RATIO=$ratio timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-legs --verify-queries 0 > gpurun_out/r2_bench7_10M_r$ratio.json 2> gpurun_out/r2_bench7_10M_r$ratio.err
done
