#!/bin/bash
# step time at small per-GPU batches with the persistent kernels on / off (which form wins when a layer has fewer items than SMs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 8 16 32; do
  echo "B=$b default:"; timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b DD_NO_PERSIST:"; DD_NO_PERSIST=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b DD_NO_PERSIST_GEMM:"; DD_NO_PERSIST_GEMM=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b DD_PS_ONE_ISSUER:"; DD_PS_ONE_ISSUER=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
  echo "B=$b DD_TC_ONE_ISSUER:"; DD_TC_ONE_ISSUER=1 timeout 300 python scripts/step_n.py $b 50 2>&1 | tail -1
done 2>&1 | tee gpurun_out/small_batch_ab.txt
