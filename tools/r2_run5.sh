#!/bin/bash
# round 2, run 5: multi-value bootstrap on the GPU + whole suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multivalue.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2e_mv.log
cat gpurun_out/r2e_mv.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2e_gpu_all.log
cat gpurun_out/r2e_gpu_all.log
