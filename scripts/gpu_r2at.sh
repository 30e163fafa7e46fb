#!/bin/bash
# round 2, pass at: dead activations' buffers reused by later layers of the same shape (the step's working set inside L2)
cd "$(dirname "$0")/.."
tag=${1:-r02_at}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_regress.py tests/test_gpu_eval.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for rep in 1 2; do for b in 64 16; do
  echo "B=$b pooled:";   timeout 300 python scripts/step_n.py $b 100 2>&1 | tail -1
  echo "B=$b one buffer per op:"; DD_NO_POOL=1 timeout 300 python scripts/step_n.py $b 100 2>&1 | tail -1
done; done
timeout 900 python -m pytest tests/test_gpu_chain_full.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "bf16_vs_reference" > gpurun_out/pytest_c_$tag.log 2>&1; echo "chain tests exit $?"; grep -E "full chain|passed|failed" gpurun_out/pytest_c_$tag.log
bash scripts/gpu_dram_step.sh $tag | tail -4
