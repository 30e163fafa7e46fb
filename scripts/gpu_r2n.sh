#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r02_n}
mkdir -p gpurun_out
DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py 1 5 7 2>&1 | tee gpurun_out/timeline_gemm_$tag.txt
