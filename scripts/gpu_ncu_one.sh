#!/bin/bash
# one `ncu --set full` capture with the finest PC-sampling interval: scripts/gpu_ncu_one.sh <kernel regex> <skip> <tag> [batch]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rx=$1; skip=$2; tag=$3; b=${4:-64}
timeout 300 python scripts/step_n.py $b 2 > gpurun_out/plain_ncu.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ncu.log; exit 1; }
timeout 600 ncu --set full --sampling-interval 0 --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_$tag -f python scripts/step_n.py $b 2 > gpurun_out/ncu_$tag.log 2>&1
echo "ncu $tag exit $?"
