#!/bin/bash
# same-box A/B of the sampling step: this build against variant libraries (gpurun_<name>_libddb200.so, scripts/build_variant.sh):
#   scripts/gpu_variant_ab.sh <tag> <variant> [<variant> ...]     -> gpurun_out/variant_ab_<tag>.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
{
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "conv or attn or unet or ln" 2>&1 | tail -2
for rep in 1 2; do
  for b in 64 8; do
    echo "B=$b this build:"; timeout 300 python scripts/step_n.py $b 100 2>&1 | tail -1
    for v in "$@"; do echo "B=$b variant $v:"; DD_LIB_PATH=$PWD/gpurun_${v}_libddb200.so timeout 300 python scripts/step_n.py $b 100 2>&1 | tail -1; done
  done
done
} 2>&1 | tee gpurun_out/variant_ab_$tag.txt
