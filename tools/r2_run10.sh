#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multivalue.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2j_tests.log; cat gpurun_out/r2j_tests.log
timeout 300 python tools/multi_value_bench.py > gpurun_out/r2j_multi_value_bench.jsonl 2> gpurun_out/r2j_mv.err; cat gpurun_out/r2j_multi_value_bench.jsonl; tail -3 gpurun_out/r2j_mv.err
