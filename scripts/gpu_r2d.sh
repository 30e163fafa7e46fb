#!/bin/bash
# round 2, pass d: where does the fused GroupNorm epilogue's time go?  (in-kernel timelines, cluster residency)
cd "$(dirname "$0")/.."
tag=${1:-r02_d}
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/clusters_$tag.txt
from downsampled_diffusion_b200 import _lib as L
import torch; torch.zeros(1, device="cuda")
for c in (1, 2, 4, 8):
    print("halo kernel, cluster of", c, "-> resident clusters:", L.lib().dd_debug_max_clusters(c))
PY
echo "== fused"; DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py 0 2 4 12 19 2>&1 | tee gpurun_out/timeline_fused_$tag.txt
echo "== unfused"; DD_NO_GN_FUSE=1 DD_NO_LN_FOLD=1 DD_LIB_PATH=$PWD/gpurun_tl_libddb200.so timeout 300 python scripts/timeline.py 0 2 4 12 19 2>&1 | tee gpurun_out/timeline_unfused_$tag.txt
