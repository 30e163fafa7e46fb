#!/bin/bash
# round 2, pass l: the bench line (sampling + training sub-record), reference arm, training launch list
cd "$(dirname "$0")/.."
tag=${1:-r02_l}
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_$tag.json 2> gpurun_out/bench_reference_$tag.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_reference_$tag.json
timeout 1200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -3 gpurun_out/bench_$tag.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "unet_step_ms", "gpu_launches")}, d["e2e"], d["default_call"], d["roofline"]["frac"], d["roofline"]["conv_ms_per_unet_step"], d.get("cpu_baseline"))
    t = d.get("train"); print("train:", t and {k: t[k] for k in ("value", "ms_per_step")}, t and t["e2e"], t and t["roofline"])
except Exception as e:
    print("parse failed", e)
PY
DD_NO_RELAYOUT_PLAN=1 timeout 600 python scripts/train_n.py 32 3 > gpurun_out/train_plain_$tag.log 2>&1 && tail -1 gpurun_out/train_plain_$tag.log && \
DD_NO_RELAYOUT_PLAN=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_$tag.csv python scripts/train_n.py 32 2 > gpurun_out/train_ncu_$tag.log 2>&1
python scripts/ncu_list.py gpurun_out/train_launches_$tag.csv > gpurun_out/train_list_$tag.txt 2>&1; head -40 gpurun_out/train_list_$tag.txt
