#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest19.log
timeout 900 python scripts/run_hotpot_synth.py > gpurun_out/r2_hotpot_synth.json 2> gpurun_out/r2_hotpot_synth.err; echo "rc=$?" >> gpurun_out/r2_hotpot_synth.err
