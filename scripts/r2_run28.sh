#!/bin/bash
# same-box A/B of the fp32 tensor kernel: pre-gang tree (3 mid buffers), HEAD (gang + 8 buffers), working tree (K-step partials)
mkdir -p gpurun_out
( cd build/ab/tree_old && timeout 600 python scripts/bench_configs.py f32_5Mx768_B16_tensor ) > gpurun_out/r2_f32_ab_old.jsonl 2> gpurun_out/r2_f32_ab_old.err
B200R_LIB=$PWD/build/ab/libb200r_headA.so timeout 600 python scripts/bench_configs.py f32_5Mx768_B16_tensor > gpurun_out/r2_f32_ab_head.jsonl 2> gpurun_out/r2_f32_ab_head.err
timeout 600 python scripts/bench_configs.py f32_5Mx768_B16_tensor > gpurun_out/r2_f32_ab_new.jsonl 2> gpurun_out/r2_f32_ab_new.err
( cd build/ab/tree_old && timeout 600 python scripts/bench_configs.py f32_5Mx768_B16_tensor ) > gpurun_out/r2_f32_ab_old2.jsonl 2> gpurun_out/r2_f32_ab_old2.err
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/r2_f32_ab_smi.txt
