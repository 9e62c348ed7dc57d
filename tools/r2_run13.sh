#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/perf_pbs.py A3 296,1184,4736 2>&1 | grep -v keygen | tee gpurun_out/r2x_perf.log
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py tests/test_gpu_cluster.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2x_parity.log
timeout 300 python tools/latency_probe.py A3 1,33 > gpurun_out/r2x_latency_A3.jsonl 2> gpurun_out/r2x_latency.err
python - <<'PY'
import json
for l in open('gpurun_out/r2x_latency_A3.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['batch'], d['cluster'], d['ms_blind_rotate'], d['failures'])
PY
