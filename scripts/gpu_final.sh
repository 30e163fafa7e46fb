#!/bin/bash
# final check of a build: the whole GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
tag=${1:-final}
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
