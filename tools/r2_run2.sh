#!/bin/bash
# round 2, run 2: tests again (memcpy race fixed), ncu source-level captures of the cluster kernel and the throughput kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log
timeout 120 python tools/probe_one.py A3 1 4 3 > gpurun_out/r2b_probe.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_blind_rotate_cl -c 1 --launch-skip 1 -f -o gpurun_out/r2b_cl4 python tools/probe_one.py A3 1 4 3 > gpurun_out/r2b_ncu_cl4.log 2>&1
tail -3 gpurun_out/r2b_ncu_cl4.log
timeout 120 python tools/probe_one.py A3 592 1 3 >> gpurun_out/r2b_probe.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_blind_rotate2 -c 1 --launch-skip 2 -f -o gpurun_out/r2b_br2 python tools/probe_one.py A3 592 1 3 > gpurun_out/r2b_ncu_br2.log 2>&1
tail -3 gpurun_out/r2b_ncu_br2.log
cat gpurun_out/r2b_probe.log
ls -la gpurun_out/*.ncu-rep
