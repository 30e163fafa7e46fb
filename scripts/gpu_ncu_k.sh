#!/bin/bash
# usage: gpu_ncu_k.sh <kernel-name-regex> <skip> <tag>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/step_n.py 64 2 > gpurun_out/plain_ncu.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$3 -f python scripts/step_n.py 64 2 > gpurun_out/ncu_$3.log 2>&1; echo "ncu exit $?"
