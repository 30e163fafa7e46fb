#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gn_mish" -p no:cacheprovider > gpurun_out/k_gn.log 2>&1; tail -3 gpurun_out/k_gn.log
timeout 300 python scripts/debug_unet.py fp32 > gpurun_out/dbg_fp32.log 2>&1; tail -12 gpurun_out/dbg_fp32.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/debug_unet.py fp32 > gpurun_out/san_fp32.log 2>&1; grep -A12 "Invalid\|ERROR SUMMARY" gpurun_out/san_fp32.log | head -60
