#!/bin/bash
# round 2 multi-GPU pack: usage tools/r2_multi.sh N [aes_batch]   (run under gpurun --gpus N)
cd "$(dirname "$0")/.."
N=${1:-2}; AESB=${2:-64}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2m_smi_${N}.txt
# BASELINE configs[1] instance-sharded + configs[3] node-sharded (the `nodes` object) on N GPUs
timeout 900 $TR --master-port 29601 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2m_bench_${N}gpu.json 2> gpurun_out/r2m_bench_${N}gpu.err
tail -c 2500 gpurun_out/r2m_bench_${N}gpu.json; tail -3 gpurun_out/r2m_bench_${N}gpu.err
# BASELINE configs[4]: PBS sweep, one sweep per GPU
timeout 600 $TR --master-port 29602 tools/pbs_sweep.py A3 1,64,296,4736,65536 > gpurun_out/r2m_pbs_sweep_${N}gpu.jsonl 2> gpurun_out/r2m_sweep_${N}gpu.err
tail -4 gpurun_out/r2m_pbs_sweep_${N}gpu.jsonl; tail -3 gpurun_out/r2m_sweep_${N}gpu.err
# BASELINE configs[2]: AES-128, fbs_size 11, instance-sharded, time-boxed
timeout 900 $TR --master-port 29603 bench.py --gpus $N --workload aes128_p11 --batch $AESB --steps 1 --warmup 1 --no-e2e --no-nodes --no-cpu-baseline > gpurun_out/r2m_aes128_b${AESB}_${N}gpu.json 2> gpurun_out/r2m_aes_${N}gpu.err
tail -c 1500 gpurun_out/r2m_aes128_b${AESB}_${N}gpu.json; tail -3 gpurun_out/r2m_aes_${N}gpu.err
# real multi-rank parity of the exchanges
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2m_multi_test_${N}gpu.log
cat gpurun_out/r2m_multi_test_${N}gpu.log
