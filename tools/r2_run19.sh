#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FBS_B200_LIB=$PWD/build_exp/libfbs_phase.so timeout 300 python tools/phase_clock.py A3 1184 2>&1 | tail -1 | tee gpurun_out/r2_phase_clock.json
FBS_B200_LIB=$PWD/build_exp/libfbs_phase.so timeout 300 python tools/phase_clock.py A2 1184 2>&1 | tail -1 | tee -a gpurun_out/r2_phase_clock.json
