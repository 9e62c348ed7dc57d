#!/bin/bash
# round 2, run 3: blocked BSK layout (one TMA per slice), cluster latency again, whole GPU suite, first full bench line with nodes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c_cluster.log
cat gpurun_out/r2c_cluster.log
timeout 300 python tools/latency_probe.py A3 1,8,16,33,64,148 > gpurun_out/r2c_latency_A3.jsonl 2> gpurun_out/r2c_latency.err
cat gpurun_out/r2c_latency_A3.jsonl
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c_gpu_all.log
cat gpurun_out/r2c_gpu_all.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 6000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
