#!/bin/bash
# 2-GPU check of the final kernels: multi-rank parity tests (fused / fused-host / NCCL exchange, CLI over two GPUs) and the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2h_multi.log
cat gpurun_out/r2h_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err
tail -c 600 gpurun_out/r2h_bench_2gpu.json; tail -2 gpurun_out/r2h_bench_2gpu.err
