#!/bin/bash
# round 2, pass ap: U-Net input as a zero-padded 64-channel activation, first ResnetBlock through the regular fused 3x3 path
cd "$(dirname "$0")/.."
tag=${1:-r02_ad}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_regress.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel + model tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for b in 64 8; do
  echo "B=$b:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
timeout 600 python scripts/op_times.py 8 > gpurun_out/op_times_${tag}_b8.txt 2>&1; tail -8 gpurun_out/op_times_${tag}_b8.txt
echo "with the im2col first block:"; for b in 64 8; do DD_FIRST_IM2COL=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_chain_full.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "bf16_vs_reference" > gpurun_out/pytest_c_$tag.log 2>&1; echo "chain tests exit $?"; grep -E "full chain|passed|failed" gpurun_out/pytest_c_$tag.log
