#!/bin/bash
#   gpurun --gpus 2 --timeout 2400 -- bash scripts/gpu_final_2gpu.sh
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_final_g2.json 2> gpurun_out/bench_final_g2.err; echo "rc=$?" >> gpurun_out/bench_final_g2.err
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_final_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final_2gpu.log
