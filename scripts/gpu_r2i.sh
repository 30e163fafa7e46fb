#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r02_i}
mkdir -p gpurun_out
echo "== timeline, half of every weight tile loaded (what a CTA pair would ingest per SM; results wrong by construction)"
DD_PS_HALF_B=1 DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py 2 12 2>&1 | tee gpurun_out/timeline_halfb_$tag.txt
echo "B=64 half-B step:"; DD_PS_HALF_B=1 timeout 300 python scripts/step_n.py 64 50 2>&1 | tail -1
