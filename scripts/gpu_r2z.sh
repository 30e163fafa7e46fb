#!/bin/bash
# round 2, pass z: second MMA issuer in the one-CTA-per-SM forms of the generic convolution kernel (8x8 / 4x4 maps, small 1x1 layers)
cd "$(dirname "$0")/.."
tag=${1:-r02_z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_regress.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel + model tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for b in 64 8; do
  echo "B=$b two issuers:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b one issuer in the generic kernel:"; DD_TC_ONE_ISSUER=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
DD_TC_ONE_ISSUER=1 timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_${tag}_one.txt 2>&1; tail -8 gpurun_out/op_times_${tag}_one.txt
timeout 600 python scripts/op_times.py 8 > gpurun_out/op_times_${tag}_b8.txt 2>&1; tail -8 gpurun_out/op_times_${tag}_b8.txt
