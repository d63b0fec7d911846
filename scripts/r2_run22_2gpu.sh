#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest22_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest22_2gpu.log
timeout 600 python scripts/run_hotpot_synth.py > gpurun_out/r2_hotpot_synth22.json 2> gpurun_out/r2_hotpot_synth22.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench22_g2.json 2> gpurun_out/r2_bench22_g2.err; echo "rc=$?" >> gpurun_out/r2_bench22_g2.err
