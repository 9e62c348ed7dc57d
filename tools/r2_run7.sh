#!/bin/bash
# round 2, run 7: prime-split cluster kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2g_cluster.log
cat gpurun_out/r2g_cluster.log
timeout 300 python tools/latency_probe.py A3 1,8,37,40,74 > gpurun_out/r2h_latency_A3.jsonl 2> gpurun_out/r2h_latency.err
grep '"cluster": 0' gpurun_out/r2h_latency_A3.jsonl; tail -3 gpurun_out/r2h_latency.err
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r2h_gpu_all.log; cat gpurun_out/r2h_gpu_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 2500 gpurun_out/r2h_bench.json
ncu --set full --clock-control none --import-source on -k regex:k_blind_rotate_cs -s 1 -c 1 -o gpurun_out/r2h_cs8 -f python tools/probe_one.py "" 1 18 3 > gpurun_out/r2h_ncu_cs8.log 2>&1; tail -2 gpurun_out/r2h_ncu_cs8.log
