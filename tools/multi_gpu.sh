#!/bin/bash
# usage: tools/multi_gpu.sh N   -- instance-sharded bench and node-sharded bench (fused + nccl exchange) on N GPUs
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_inst_$N.json 2> gpurun_out/bench_inst_$N.err
tail -2 gpurun_out/bench_inst_$N.err; cat gpurun_out/bench_inst_$N.json | cut -c1-400
for ex in fused nccl; do
$TR --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --shard nodes --exchange $ex --workload mult16_p17 --batch 64 > gpurun_out/bench_nodes_${ex}_$N.json 2> gpurun_out/bench_nodes_${ex}_$N.err
tail -2 gpurun_out/bench_nodes_${ex}_$N.err; cat gpurun_out/bench_nodes_${ex}_$N.json
done
python bench.py --gpus 1 --steps 2 --warmup 1 --shard nodes --workload mult16_p17 --batch 64 > gpurun_out/bench_nodes_1.json 2> gpurun_out/bench_nodes_1.err
tail -2 gpurun_out/bench_nodes_1.err; cat gpurun_out/bench_nodes_1.json
