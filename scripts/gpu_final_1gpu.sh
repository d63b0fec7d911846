#!/bin/bash
# final single-GPU record of a round: full GPU suite, smoke(), the contract bench (both arms), launch list under ncu, other configs
#   gpurun --timeout 3000 -- bash scripts/gpu_final_1gpu.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_final.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "rc=$?" >> gpurun_out/bench_final_ref.err
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?" >> gpurun_out/bench_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final_10M.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-legs --verify-queries 0 > gpurun_out/ncu_launches_final.log 2>&1
timeout 900 python scripts/bench_configs.py cfg1_1Mx768_B1024 cfg2_10Mx768_B1_gemm cfg2_10Mx768_B16_gemm f32_5Mx768_B1_tensor f32_5Mx768_B16_tensor i8_10Mx768_B16_tensor i8_10Mx768_B256_tensor cfg4_graph_5M_50M_B4096 > gpurun_out/other_configs_final.jsonl 2> gpurun_out/other_configs_final.err
timeout 600 python scripts/run_hotpot_synth.py > gpurun_out/hotpot_synth_final.json 2> gpurun_out/hotpot_synth_final.err
