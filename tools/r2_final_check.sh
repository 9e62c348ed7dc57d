#!/bin/bash
# last check of HEAD on one GPU: the driver's sequence (GPU tests, smoke, both bench arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2q_gpu_all.log; cat gpurun_out/r2q_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; tail -2 gpurun_out/r2q_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 1500 gpurun_out/r2q_bench.json; tail -2 gpurun_out/r2q_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err; tail -c 600 gpurun_out/r2q_bench_reference.json
