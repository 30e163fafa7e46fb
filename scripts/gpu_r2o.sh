#!/bin/bash
# round 2, pass o: layer parameters staged once per CTA in both persistent kernels
cd "$(dirname "$0")/.."
tag=${1:-r02_o}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for b in 64 8; do
  echo "B=$b default:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b no persistent GEMM:"; DD_NO_PERSIST_GEMM=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b no persistent at all:"; DD_NO_PERSIST_GEMM=1 DD_NO_PERSIST=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_kernels.py > gpurun_out/pytest_all_$tag.log 2>&1; echo "all other gpu tests exit $?"; tail -3 gpurun_out/pytest_all_$tag.log
