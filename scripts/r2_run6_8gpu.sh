#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_comm.py -m gpu -q -k two_processes > gpurun_out/r2_pytest6_ipc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6_ipc.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench6_g$n.json 2> gpurun_out/r2_bench6_g$n.err; echo "rc=$?" >> gpurun_out/r2_bench6_g$n.err
done
