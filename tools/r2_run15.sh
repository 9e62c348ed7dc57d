#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2x_variants.log
bash tools/r2_run12.sh
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2x_parity.log
