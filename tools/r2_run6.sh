#!/bin/bash
# round 2, run 6: multi-value again, AES on A3, boot_cost table, profile pack of the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multivalue.py tests/test_gpu_aes.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2f_mv.log
cat gpurun_out/r2f_mv.log
timeout 600 python tools/boot_cost.py > gpurun_out/r2f_boot_cost.log 2>&1
tail -6 gpurun_out/r2f_boot_cost.log
timeout 1500 bash tools/profile_all.sh r2v1 2>&1 | tail -8
ls -la gpurun_out/r2v1_*
