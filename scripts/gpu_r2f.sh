#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r02_f}
mkdir -p gpurun_out
echo "== persistent"; DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py 2 3 4 12 2>&1 | tee gpurun_out/timeline_persist_$tag.txt
