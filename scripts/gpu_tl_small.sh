#!/bin/bash
# in-kernel timelines of the 8x8 / 4x4 layers (instrumented build gpurun_tl_libddb200.so): scripts/gpu_tl_small.sh [conv indices]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py ${@:-16 17 23 24} 2>&1 | tee gpurun_out/timeline_small.txt
