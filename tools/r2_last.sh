#!/bin/bash
# last check of the round: all GPU tests, smoke, both bench arms (what the driver runs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2w_gpu_all.log; cat gpurun_out/r2w_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; tail -2 gpurun_out/r2w_smoke.log
timeout 600 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -c 400 gpurun_out/r2w_bench.json; tail -2 gpurun_out/r2w_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; tail -c 200 gpurun_out/r2w_bench_reference.json
