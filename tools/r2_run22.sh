#!/bin/bash
# last check of the round + profile pack of the final binaries
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/r2_last.sh
timeout 600 bash tools/profile_all.sh r2v7
