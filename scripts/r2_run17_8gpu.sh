#!/bin/bash
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench17_g$n.json 2> gpurun_out/r2_bench17_g$n.err; echo "rc=$?" >> gpurun_out/r2_bench17_g$n.err
done
timeout 300 python -m pytest tests/test_gpu_comm.py tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest17.log
timeout 300 python -m pytest tests/test_gpu_backend.py -m gpu -q > gpurun_out/r2_pytest17b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest17b.log
