#!/bin/bash
# A/B of blind-rotation kernel variants built into build_exp/*.so (FBS_B200_LIB selects the library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" $(ls build_exp/libfbs_*.so 2>/dev/null); do
  echo "== ${lib:-default}" | tee -a gpurun_out/r2x_variants.log
  FBS_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tools/perf_pbs.py A3 296,1184 2>&1 | grep -v keygen | tee -a gpurun_out/r2x_variants.log
done
