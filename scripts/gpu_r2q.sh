#!/bin/bash
# round 2, pass q: what the MMA warp of the persistent convolution spends its time on -- per-tap tcgen05 fence and the
# position of the mbarrier waits (look-ahead) as compile-time variants (scripts/build_variant.sh)
cd "$(dirname "$0")/.."
tag=${1:-r02_q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; fi
for v in la0_f1 la0_f0 la1_f0 la2_f1 la2_f0 la3_f0; do
  echo "== timeline $v"; DD_LIB_PATH=$PWD/gpurun_tl_${v}_libddb200.so timeout 300 python scripts/timeline.py 2 12 2>&1 | tee -a gpurun_out/timeline_variants_$tag.txt
done
for b in 64 8; do
  echo "B=$b default (look-ahead 2, fence):";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  for v in la0_f1 la0_f0 la2_f0; do
    echo "B=$b $v:"; DD_LIB_PATH=$PWD/gpurun_${v}_libddb200.so timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  done
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
