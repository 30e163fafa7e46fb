#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r02_j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "gn_fused" > gpurun_out/pytest_k_$tag.log 2>&1; echo "x1 kernel tests exit $?"; tail -2 gpurun_out/pytest_k_$tag.log
DD_PERSIST_TWO=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "gn_fused" > gpurun_out/pytest_k2_$tag.log 2>&1; echo "x2 kernel tests exit $?"; tail -2 gpurun_out/pytest_k2_$tag.log
for b in 64 8; do
  echo "B=$b persistent x1:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b persistent x2:";   DD_TC_VERBOSE=1 DD_PERSIST_TWO=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | grep -E "step ms|occupancy" | sort | uniq | tail -3
done
DD_PERSIST_TWO=1 timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_x2_$tag.txt 2>&1; head -20 gpurun_out/op_times_x2_$tag.txt | tail -16
