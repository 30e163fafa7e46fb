#!/bin/bash
# Round-end pass without the ncu captures (the sampling kernels did not change since the last full pass): parity suite, both
# bench arms, training bench, the SURVEY 8(f) row timings.  Usage: bash scripts/gpu_final_lite.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r01_g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err; echo "ref exit $?"
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_$tag.json
timeout 900 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train_$tag.json 2> gpurun_out/bench_train_$tag.err; echo "train exit $?"; cut -c1-200 gpurun_out/bench_train_$tag.json
timeout 600 python scripts/f_rows_bench.py --cpu > gpurun_out/f_rows_$tag.json 2> gpurun_out/f_rows_$tag.err; echo "f rows exit $?"
