#!/bin/bash
#   gpurun --gpus 8 --timeout 1500 -- bash scripts/gpu_final_8gpu.sh
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_final_g$n.json 2> gpurun_out/bench_final_g$n.err; echo "rc=$?" >> gpurun_out/bench_final_g$n.err
done
timeout 400 python -m pytest tests/test_gpu_comm.py tests/test_gpu_sharded.py tests/test_gpu_backend.py -m gpu -q > gpurun_out/pytest_final_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final_8gpu.log
