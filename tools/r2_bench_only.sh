#!/bin/bash
# bench line only on N GPUs: tools/r2_bench_only.sh N
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2p_bench_${N}gpu.json 2> gpurun_out/r2p_bench_${N}gpu.err
tail -c 2500 gpurun_out/r2p_bench_${N}gpu.json; tail -3 gpurun_out/r2p_bench_${N}gpu.err
