#!/bin/bash
# round 2, pass h: software-pipelined persistent epilogue, ftz Mish
cd "$(dirname "$0")/.."
tag=${1:-r02_h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_k_$tag.log 2>&1; rc=$?; echo "kernel tests exit $rc"; tail -3 gpurun_out/pytest_k_$tag.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|error" gpurun_out/pytest_k_$tag.log | head -20; exit 0; fi
for b in 64 8; do
  echo "B=$b persistent x1:";   timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b persistent x2:";   DD_TC_VERBOSE=1 DD_PERSIST_TWO=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | grep -E "step ms|occupancy" | sort | uniq | tail -3
  echo "B=$b no persistent:"; DD_NO_PERSIST=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b unfused:"; DD_NO_GN_FUSE=1 DD_NO_LN_FOLD=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done
timeout 600 python scripts/op_times.py 64 > gpurun_out/op_times_$tag.txt 2>&1; tail -8 gpurun_out/op_times_$tag.txt
echo "== timeline"; DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py 2 4 12 2>&1 | tee gpurun_out/timeline_persist_$tag.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_regress.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_m_$tag.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/pytest_m_$tag.log
timeout 900 python -m pytest tests/test_gpu_chain_full.py -m gpu -q -s --timeout 600 -p no:cacheprovider -k "bf16_vs_reference" > gpurun_out/pytest_c_$tag.log 2>&1; echo "chain tests exit $?"; grep -E "full chain|passed|failed" gpurun_out/pytest_c_$tag.log
