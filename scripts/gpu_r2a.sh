#!/bin/bash
# round 2, pass a: tcgen05.mma issue-rate micro-benchmark, the new full-chain / regression tests, then the whole GPU suite
cd "$(dirname "$0")/.."
tag=${1:-r02_a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
timeout 120 ./scripts/micro/pair_mma 2000 > gpurun_out/pair_mma_$tag.txt 2>&1; echo "pair_mma exit $?"; cat gpurun_out/pair_mma_$tag.txt
timeout 1500 python -m pytest tests/test_gpu_chain_full.py tests/test_gpu_regress.py -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_new_$tag.log 2>&1; echo "new tests exit $?"
grep -E "full chain|C4-size|passed|failed|Error|error|assert" gpurun_out/pytest_new_$tag.log | tail -40
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_chain_full.py --deselect tests/test_gpu_regress.py > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
