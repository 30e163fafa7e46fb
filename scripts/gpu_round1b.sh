#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
timeout 300 python scripts/debug_unet.py fp32 2>&1 | tail -4
timeout 300 python scripts/debug_unet.py bf16 2>&1 | tail -4
run m_fp32 tests/test_gpu_model.py -m gpu -k "fp32 or rng or convolutional or roundtrip or rejects" -s
run m_bf16 tests/test_gpu_model.py -m gpu -k "bf16 and not rejects" -s
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
cat gpurun_out/summary.txt
