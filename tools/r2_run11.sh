#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2k_cluster.log; cat gpurun_out/r2k_cluster.log
timeout 300 python tools/latency_probe.py A3 1,16,33,64 > gpurun_out/r2k_latency_A3.jsonl 2> gpurun_out/r2k_latency.err
python - <<'PY'
import json
for l in open('gpurun_out/r2k_latency_A3.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['batch'], d['cluster'], d['ms_blind_rotate'], d['failures'])
PY
tail -3 gpurun_out/r2k_latency.err
