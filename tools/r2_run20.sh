#!/bin/bash
# refresh of the one-GPU side measurements with the final kernels: AES-128 (configs[2]), multi-value bootstrap, config-1 line, launch overhead
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --workload aes128_p11 --batch 64 --steps 1 --warmup 1 --no-e2e --no-nodes --no-cpu-baseline > gpurun_out/r2i_aes128_b64_1gpu.json 2> gpurun_out/r2i_aes.err; tail -c 500 gpurun_out/r2i_aes128_b64_1gpu.json
timeout 300 python tools/multi_value_bench.py > gpurun_out/r2i_multi_value_bench.jsonl 2> gpurun_out/r2i_mv.err; cat gpurun_out/r2i_multi_value_bench.jsonl | cut -c1-400; tail -2 gpurun_out/r2i_mv.err
timeout 300 python bench.py --config1 > gpurun_out/r2i_bench_config1.json 2> gpurun_out/r2i_config1.err; tail -c 300 gpurun_out/r2i_bench_config1.json
