#!/bin/bash
# final 2-GPU checks: multi-rank parity tests (incl. the CLI over two GPUs), the PBS sweep on 2 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2y_multi.log
cat gpurun_out/r2y_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/pbs_sweep.py A3 1,64,296,4736,65536 > gpurun_out/r2y_pbs_sweep_2gpu.jsonl 2> gpurun_out/r2y_sweep.err
grep "^{" gpurun_out/r2y_pbs_sweep_2gpu.jsonl | tail -3; tail -2 gpurun_out/r2y_sweep.err
