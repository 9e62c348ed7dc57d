#!/bin/bash
# 8-GPU bench line with the final kernels (instances + node-sharded 64x64 multiplier)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2j_bench_8gpu.json 2> gpurun_out/r2j_bench_8gpu.err
tail -c 700 gpurun_out/r2j_bench_8gpu.json; tail -2 gpurun_out/r2j_bench_8gpu.err
