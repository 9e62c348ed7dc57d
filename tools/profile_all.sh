#!/bin/bash
# Round profile pack: plain bench, launch list (per-kernel share of a step), full captures of the hot kernels.
# usage: tools/profile_all.sh [tag]   (writes gpurun_out/<tag>_*)
cd "$(dirname "$0")/.."
TAG=${1:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 296 --no-cpu-baseline --no-e2e --no-nodes"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
python tools/perf_pbs.py "" 592 > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_blind_rotate|k_keyswitch_mma|k_lincomb" -s 3 -c 3 -o gpurun_out/${TAG}_hot -f python tools/perf_pbs.py "" 592 > gpurun_out/${TAG}_ncu_hot.log 2>&1
# the cluster-split low-latency kernel (one bootstrap over 4 CTAs)
python tools/probe_one.py "" 1 4 3 > gpurun_out/${TAG}_cl_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_blind_rotate_cl -s 1 -c 1 -o gpurun_out/${TAG}_cl4 -f python tools/probe_one.py "" 1 4 3 > gpurun_out/${TAG}_ncu_cl4.log 2>&1
tail -2 gpurun_out/${TAG}_plain2.log; tail -2 gpurun_out/${TAG}_ncu_hot.log; wc -l gpurun_out/${TAG}_launches.csv
