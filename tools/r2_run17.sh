#!/bin/bash
# final-state check of the round: all GPU tests, smoke, both bench arms, config-1 line, the profile pack (launch list + ncu captures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2g_gpu_all.log; cat gpurun_out/r2g_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -2 gpurun_out/r2g_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 300 gpurun_out/r2g_bench.json; tail -2 gpurun_out/r2g_bench.err
timeout 600 python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; tail -c 200 gpurun_out/r2g_bench_default.json
timeout 600 python bench.py --impl reference > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err; tail -c 300 gpurun_out/r2g_bench_reference.json
timeout 900 bash tools/profile_all.sh r2v7
timeout 600 python tools/boot_cost.py > gpurun_out/r2g_boot_cost.log 2>&1; tail -3 gpurun_out/r2g_boot_cost.log | cut -c1-200
timeout 600 python tools/pbs_sweep.py > gpurun_out/r2g_pbs_sweep_1gpu.jsonl 2> gpurun_out/r2g_sweep.err; tail -2 gpurun_out/r2g_pbs_sweep_1gpu.jsonl | cut -c1-300
