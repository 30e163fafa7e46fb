#!/bin/bash
# Full GPU pass: parity suite, smoke, bench.  Usage: bash scripts/gpu_all.sh [tag]
cd "$(dirname "$0")/.."
tag=${1:-run}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
