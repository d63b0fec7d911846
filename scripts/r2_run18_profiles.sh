#!/bin/bash
mkdir -p gpurun_out
# launch list of the contract command (short run), then --set full of the dominant K1 launch (last phase) and of the fp32 tensor kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_10M.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-legs --verify-queries 0 > gpurun_out/r2_ncu_l10.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 11 -c 1 -o gpurun_out/r2_k1_main -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-legs --no-e2e --verify-queries 0 > gpurun_out/r2_ncu_k1main.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_f32 -s 5 -c 1 -o gpurun_out/r2_rows_f32_tmem -f \
  python scripts/bench_configs.py f32_5Mx768_B16_tensor > gpurun_out/r2_ncu_rows_f32b.log 2>&1
timeout 900 python scripts/bench_configs.py cfg1_1Mx768_B1024 f32_5Mx768_B1_tensor f32_5Mx768_B8_tensor f32_5Mx768_B16_tensor f32_5Mx768_B16_scan i8_10Mx768_B16_tensor cfg4_graph_5M_50M_B4096 > gpurun_out/r2_other_configs.jsonl 2> gpurun_out/r2_other_configs.err
