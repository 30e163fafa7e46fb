#!/bin/bash
# round 2, pass ar: the 1x1 res_conv of a ResnetBlock on a side stream (a parallel branch of the captured graph)
cd "$(dirname "$0")/.."
tag=${1:-r02_ar}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_regress.py tests/test_gpu_eval.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for b in 64 32 8; do
  echo "B=$b fork:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b in line:"; DD_NO_FORK=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_chain_full.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "bf16_vs_reference" > gpurun_out/pytest_c_$tag.log 2>&1; echo "chain tests exit $?"; grep -E "full chain|passed|failed" gpurun_out/pytest_c_$tag.log
