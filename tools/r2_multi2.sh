#!/bin/bash
# round 2: AES-128 at 256 instances per GPU (BASELINE configs[2]) and the round-1 node-sharded workload (mult16) on N GPUs
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2n_parity.log; cat gpurun_out/r2n_parity.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29611 bench.py --gpus $N --workload aes128_p11 --batch 256 --steps 1 --warmup 1 --no-e2e --no-nodes --no-cpu-baseline > gpurun_out/r2n_aes128_b256_${N}gpu.json 2> gpurun_out/r2n_aes_${N}gpu.err
tail -c 1200 gpurun_out/r2n_aes128_b256_${N}gpu.json; tail -3 gpurun_out/r2n_aes_${N}gpu.err
for B in 1 64; do
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 3 --warmup 1 --shard nodes --workload mult16_p17 --batch $B > gpurun_out/r2n_nodes_mult16_b${B}_${N}gpu.json 2> gpurun_out/r2n_nodes_${N}gpu.err
tail -c 900 gpurun_out/r2n_nodes_mult16_b${B}_${N}gpu.json; tail -2 gpurun_out/r2n_nodes_${N}gpu.err
done
