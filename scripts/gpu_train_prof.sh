#!/bin/bash
# training step: plain timing, then an ncu launch list (gpu__time_duration) of two steps
cd "$(dirname "$0")/.."
tag=${1:-t}
mkdir -p gpurun_out
timeout 600 python scripts/train_n.py 32 3 > gpurun_out/train_plain_$tag.log 2>&1; tail -1 gpurun_out/train_plain_$tag.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_$tag.csv python scripts/train_n.py 32 2 > gpurun_out/train_ncu_$tag.log 2>&1
python scripts/ncu_list.py gpurun_out/train_launches_$tag.csv x > gpurun_out/train_list_$tag.txt 2>&1; head -30 gpurun_out/train_list_$tag.txt
