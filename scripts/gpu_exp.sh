#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "0 0" "3 0" "3 1" "5 0" "5 1"; do
  set -- $cfg
  DD_EXP_SHIFT=$1 DD_EXP_BO=$2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_paths and bf16" -p no:cacheprovider > gpurun_out/exp_$1_$2.log 2>&1
  echo "shift=$1 base_offset=$2 -> exit $? : $(tail -1 gpurun_out/exp_$1_$2.log)"
done
