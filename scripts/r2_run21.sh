#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log
# the N=8 shard size on one GPU: what the per-phase selects and the phase schedule cost next to 5.4 ms of tile kernels
for r in 8 16 32; do
  B200R_PHASE_RATIO=$r timeout 600 python bench.py --rows 1250000 --steps 20 --warmup 5 --no-cpu-baseline --no-legs --verify-queries 4 > gpurun_out/r2_bench21_shard_ratio$r.json 2> gpurun_out/r2_bench21_shard_ratio$r.err
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-legs > gpurun_out/r2_bench21_10M.json 2> gpurun_out/r2_bench21_10M.err
timeout 600 python scripts/run_hotpot_synth.py > gpurun_out/r2_hotpot_synth21.json 2> gpurun_out/r2_hotpot_synth21.err
