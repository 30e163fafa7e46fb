#!/bin/bash
# in-kernel timelines of the persistent convolution from the instrumented build(s): scripts/gpu_tl.sh <tag> [variant ...]
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "== timeline $v"; DD_LIB_PATH=$PWD/gpurun_${v}_libddb200.so timeout 300 python scripts/timeline.py 2 12 2>&1 | tee -a gpurun_out/timeline_variants_$tag.txt
done
