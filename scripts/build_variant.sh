#!/bin/bash
# Builds an instrumented / experimental copy of libddb200.so at the repo root: gpurun_<name>_libddb200.so
# (git-ignored, travels with gpurun; load it with DD_LIB_PATH).   usage: scripts/build_variant.sh tl -DDD_TC_TIMELINE=1 [...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
d=$(mktemp -d)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
objs=""
for f in downsampled_diffusion_b200/csrc/*.cu; do
  o="$d/$(basename "${f%.cu}").o"
  $NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c "$f" -o "$o" &
  objs="$objs $o"
done
wait
$NVCC $ARCH -shared -o "gpurun_${name}_libddb200.so" $objs
rm -rf "$d"
echo "built gpurun_${name}_libddb200.so"
