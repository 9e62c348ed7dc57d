#!/bin/bash
# round 2, first GPU bring-up of the cluster-split kernel + device-side hand-off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_cluster.log
cat gpurun_out/r2_cluster.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "handoff or fused" 2>&1 | tail -40 > gpurun_out/r2_handoff.log
cat gpurun_out/r2_handoff.log
timeout 300 python tools/latency_probe.py A3 > gpurun_out/r2_latency_A3.jsonl 2> gpurun_out/r2_latency.err
cat gpurun_out/r2_latency_A3.jsonl | tail -40
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_gpu_all.log
cat gpurun_out/r2_gpu_all.log
