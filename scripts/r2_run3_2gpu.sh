#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r2_pytest3_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench3_g2.json 2> gpurun_out/r2_bench3_g2.err; echo "rc=$?" >> gpurun_out/r2_bench3_g2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl --verify-queries 0 > gpurun_out/r2_bench3_g2_nccl.json 2> gpurun_out/r2_bench3_g2_nccl.err
# ncu of one per-phase select (threshold mode) and the final select at the N=8 shard size
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_lists -s 7 -c 3 -o gpurun_out/r2_select -f \
  python bench.py --rows 1250000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --verify-queries 0 > gpurun_out/r2_ncu_select.log 2>&1
