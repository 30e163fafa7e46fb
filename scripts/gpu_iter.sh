#!/bin/bash
# Iteration pass: parity suite, step timing, ncu launch list of 3 steps.  Usage: bash scripts/gpu_iter.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-it}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_$tag.log
timeout 300 python scripts/step_n.py 64 50 > gpurun_out/plain_step_$tag.log 2>&1; tail -1 gpurun_out/plain_step_$tag.log
timeout 300 python scripts/step_n.py 64 3 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python scripts/step_n.py 64 3 > gpurun_out/ncu_step_$tag.log 2>&1
python scripts/ncu_list.py gpurun_out/launches_$tag.csv x > gpurun_out/list_$tag.txt 2>&1; head -14 gpurun_out/list_$tag.txt
