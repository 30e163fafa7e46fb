#!/bin/bash
# round 2, pass p: halo boxes of the persistent convolution requested by their own producer warp (A/B against the in-order producer)
cd "$(dirname "$0")/.."
tag=${1:-r02_p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for b in 64 8; do
  echo "B=$b halo warp:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b in-order producer:"; DD_LIB_PATH=$PWD/gpurun_nohw_libddb200.so timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
echo "== timeline, halo warp"; DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py 2 4 12 2>&1 | tee gpurun_out/timeline_persist_$tag.txt
echo "== timeline, in-order producer"; DD_LIB_PATH=$PWD/gpurun_tl0_libddb200.so timeout 300 python scripts/timeline.py 2 4 12 2>&1 | tee gpurun_out/timeline_persist0_$tag.txt
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
