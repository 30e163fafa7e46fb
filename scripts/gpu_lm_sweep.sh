#!/bin/bash
# linear-attention context kernel: CTAs aimed at per launch (splits of the pixel range) against the per-op times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 64 8; do for t in 148 296 444 600 4096; do
  echo "B=$b target $t:"; DD_LM_TARGET_CTAS=$t timeout 300 python scripts/op_times.py $b 2>&1 | grep -E "dd_linattn_mix  |linattn_mix +n=" | awk '{printf "%s ", $3} END {print ""}'
done; done 2>&1 | tee gpurun_out/lm_sweep.txt
