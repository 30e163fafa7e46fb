#!/bin/bash
# First GPU pass: each group in its own process so a trapped kernel does not poison later groups.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 300 -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
rm -f gpurun_out/summary.txt
run k_elem tests/test_gpu_kernels.py -m gpu -k "not conv_paths and not conv_tc"
run k_conv_fp32 tests/test_gpu_kernels.py -m gpu -k "conv_paths and fp32"
run k_conv_bf16 tests/test_gpu_kernels.py -m gpu -k "conv_paths and bf16 or conv_tc"
run m_fp32 tests/test_gpu_model.py -m gpu -k "fp32 or rng or convolutional or roundtrip or rejects" -s
run m_bf16 tests/test_gpu_model.py -m gpu -k "bf16 and not rejects" -s
cat gpurun_out/summary.txt
