#!/bin/bash
# usage: gpu_ncu_multi.sh "<regex>:<skip>:<tag>" ...   -- one `ncu --set full` capture per argument (after a plain run exits 0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/step_n.py 64 2 > gpurun_out/plain_ncu.log 2>&1 || { echo "plain run failed"; exit 1; }
for spec in "$@"; do
  IFS=: read -r rx skip tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/prof_$tag -f python scripts/step_n.py 64 2 > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu $tag exit $?"
done
