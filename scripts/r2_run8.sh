#!/bin/bash
mkdir -p gpurun_out
for ratio in 3 4 6 8; do
for w in 1 0; do
B200R_SELECT_WARP=$w B200R_PHASE_RATIO=$ratio timeout 300 python bench.py --rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline --no-legs --no-e2e --verify-queries 2 > gpurun_out/r2_bench8_r${ratio}_w$w.json 2> gpurun_out/r2_bench8_r${ratio}_w$w.err
done
done
