#!/bin/bash
# round 2, run 7: prime-split cluster kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2g_cluster.log
cat gpurun_out/r2g_cluster.log
timeout 300 python tools/latency_probe.py A3 1,16,33,64 > gpurun_out/r2g_latency_A3.jsonl 2> gpurun_out/r2g_latency.err
cat gpurun_out/r2g_latency_A3.jsonl; tail -3 gpurun_out/r2g_latency.err
