#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/boot_cost.py > gpurun_out/r2k_boot_cost.log 2>&1; tail -2 gpurun_out/r2k_boot_cost.log | cut -c1-200
