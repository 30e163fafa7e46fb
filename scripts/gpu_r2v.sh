#!/bin/bash
# round 2, pass v: GroupNorm partial sums exchanged as flagged 16-byte packets (no atomics, no gpu-scope fence)
cd "$(dirname "$0")/.."
tag=${1:-r02_v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for v in tl tl_ew8; do
  echo "== timeline $v"; DD_LIB_PATH=$PWD/gpurun_${v}_libddb200.so timeout 300 python scripts/timeline.py 2 12 2>&1 | tee -a gpurun_out/timeline_variants_$tag.txt
done
for b in 64 8; do
  echo "B=$b 2 MMA warps, 16 epilogue warps:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  for v in mw1 ew8 mw1_ew8; do
  echo "B=$b $v:"; DD_LIB_PATH=$PWD/gpurun_${v}_libddb200.so timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  done
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_m_$tag.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/pytest_m_$tag.log
