#!/bin/bash
# round 2, pass ag: weight producer of the generic kernel ahead of the grid dependency too
cd "$(dirname "$0")/.."
tag=${1:-r02_ag}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_regress.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel + model tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi

for b in 64 8; do
  echo "B=$b:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
