#!/bin/bash
# round 2, run 4 (2 GPUs): real multi-rank parity test of the node-sharded exchanges, bench line at N=2 with the nodes object
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2d_smi.txt
timeout 300 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2d_cluster.log
cat gpurun_out/r2d_cluster.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2d_multi.log
cat gpurun_out/r2d_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err
tail -c 5000 gpurun_out/r2d_bench_2gpu.json; tail -5 gpurun_out/r2d_bench_2gpu.err
for ex in fused-host nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 3 --warmup 1 --shard nodes --workload mult64_p17 --batch 1 --exchange $ex > gpurun_out/r2d_nodes_b1_${ex}_2gpu.json 2> gpurun_out/r2d_nodes_${ex}.err
tail -c 1500 gpurun_out/r2d_nodes_b1_${ex}_2gpu.json; tail -3 gpurun_out/r2d_nodes_${ex}.err
done
