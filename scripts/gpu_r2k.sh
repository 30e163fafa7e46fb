#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r02_k}
mkdir -p gpurun_out
DD_PERSIST_TWO=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider -k "gn_fused" > gpurun_out/pytest_k2_$tag.log 2>&1; echo "x2 kernel tests exit $?"; tail -2 gpurun_out/pytest_k2_$tag.log
for b in 64 8; do
  echo "B=$b persistent x2:";   DD_TC_VERBOSE=1 DD_PERSIST_TWO=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | grep -E "step ms|occupancy" | sort | uniq | tail -3
done
DD_PERSIST_TWO=1 timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_x2_$tag.txt 2>&1; head -20 gpurun_out/op_times_x2_$tag.txt | tail -16
for s in 100 250 400; do echo "stagger $s clk per k-block, no persistent:"; DD_NO_PERSIST=1 DD_TC_STAGGER=$s timeout 300 python scripts/step_n.py 64 50 2>&1 | tail -1; done
echo "stagger 250, unfused:"; DD_NO_GN_FUSE=1 DD_NO_LN_FOLD=1 DD_TC_STAGGER=250 timeout 300 python scripts/step_n.py 64 50 2>&1 | tail -1
