#!/bin/bash
# Training iteration pass: training parity tests, then the C4 bench line with graphs on and off.  Usage: bash scripts/gpu_train_iter.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-tr}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train_$tag.json 2> gpurun_out/bench_train_$tag.err; echo "bench exit $?"
python - <<PY
import json
for f in ("gpurun_out/bench_train_$tag.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-2000:])
PY
DD_TRAIN_GRAPH=0 timeout 600 python bench.py --workload train --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager lists: ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])"
