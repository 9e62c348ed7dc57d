#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/perf_pbs.py "" 592 > gpurun_out/r2v4_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_blind_rotate2" -s 1 -c 1 -o gpurun_out/r2v4_hot -f python tools/perf_pbs.py "" 592 > gpurun_out/r2v4_ncu_hot.log 2>&1
tail -2 gpurun_out/r2v4_plain2.log; tail -2 gpurun_out/r2v4_ncu_hot.log
