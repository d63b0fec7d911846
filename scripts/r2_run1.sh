#!/bin/bash
# round-2 GPU call 1: parity suite, contract bench with the recall leg, N=8-shard-sized bench, launch list, ncu of K1 at B=16 / B=1024
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "rc=$?" >> gpurun_out/r2_bench1.err
python bench.py --rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench1_1p25M.json 2> gpurun_out/r2_bench1_1p25M.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_1p25M.csv \
  python bench.py --rows 1250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --verify-queries 0 > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_topk -c 4 -o gpurun_out/r2_k1_b16 -f \
  python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --verify-queries 0 > gpurun_out/r2_ncu_b16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_topk -c 6 -o gpurun_out/r2_k1_b1024 -f \
  python bench.py --rows 1000000 --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --verify-queries 0 > gpurun_out/r2_ncu_b1024.log 2>&1
ls -la gpurun_out/ | tail -12
