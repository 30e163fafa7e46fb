#!/bin/bash
# quick check of a kernel change: kernel + model parity tests, step time at 64 and 8 samples (three repeats: one box repeats to 0.05 %)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2 3; do for b in 64 8; do echo "B=$b:"; timeout 300 python scripts/step_n.py $b 100 2>&1 | tail -1; done; done
