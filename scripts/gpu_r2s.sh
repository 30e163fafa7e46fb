#!/bin/bash
# round 2, pass s: persistent convolution with 16 epilogue warps (32 channels of a pixel per thread) and a statistics warp
cd "$(dirname "$0")/.."
tag=${1:-r02_s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for v in tl tl_ew8; do
  echo "== timeline $v"; DD_LIB_PATH=$PWD/gpurun_${v}_libddb200.so timeout 300 python scripts/timeline.py 2 12 2>&1 | tee -a gpurun_out/timeline_variants_$tag.txt
done
for b in 64 8; do
  echo "B=$b 16 epilogue warps:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b 8 epilogue warps:"; DD_LIB_PATH=$PWD/gpurun_ew8_libddb200.so timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_m_$tag.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/pytest_m_$tag.log
