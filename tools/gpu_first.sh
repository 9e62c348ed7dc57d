#!/bin/bash
# first GPU bring-up: parity tests on toy sets, then full-size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/parity.log
cat gpurun_out/parity.log
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/full.log
cat gpurun_out/full.log
