#!/bin/bash
# round 2, final one-GPU validation: what the driver runs (tests, smoke, bench both arms) + the remaining one-GPU measurements
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2z_gpu_all.log
cat gpurun_out/r2z_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -c 3000 gpurun_out/r2z_bench.json; tail -3 gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
tail -c 1200 gpurun_out/r2z_bench_reference.json; tail -3 gpurun_out/r2z_bench_reference.err
timeout 300 python bench.py --config1 > gpurun_out/r2z_bench_config1.json 2> gpurun_out/r2z_config1.err
tail -c 1500 gpurun_out/r2z_bench_config1.json
timeout 120 python tools/launch_overhead.py > gpurun_out/r2z_launch_overhead.jsonl 2>&1; cat gpurun_out/r2z_launch_overhead.jsonl
# A/B: twiddle table in shared memory for the one-bootstrap tail kernel at M = 3
timeout 200 python tools/latency_probe.py A3 100,148 > gpurun_out/r2z_tail_base.jsonl 2>&1; grep '"cluster": 1' gpurun_out/r2z_tail_base.jsonl
FBS_B200_LIB=build_exp/libfbs_tw1.so timeout 200 python tools/latency_probe.py A3 100,148 > gpurun_out/r2z_tail_tw1.jsonl 2>&1; grep '"cluster": 1' gpurun_out/r2z_tail_tw1.jsonl
timeout 600 python tools/pbs_sweep.py A3 1,8,33,64,148,296,1184,4736,16384,65536 > gpurun_out/r2z_pbs_sweep_1gpu.jsonl 2> gpurun_out/r2z_sweep.err
tail -3 gpurun_out/r2z_pbs_sweep_1gpu.jsonl
timeout 600 python bench.py --workload aes128_p11 --batch 64 --steps 1 --warmup 1 --no-e2e --no-nodes --no-cpu-baseline > gpurun_out/r2z_aes128_b64_1gpu.json 2> gpurun_out/r2z_aes.err
tail -c 800 gpurun_out/r2z_aes128_b64_1gpu.json
