#!/bin/bash
# Sampling bench at other per-GPU batch sizes (the literal C3 split of 64 samples over 8/4/2 GPUs is 8/16/32 per GPU; 128/256 show
# where the small feature maps saturate).  Usage: bash scripts/batch_sweep.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-sweep}
mkdir -p gpurun_out
: > gpurun_out/batch_sweep_$tag.jsonl
for b in 8 16 32 128 256; do
  timeout 600 python bench.py --batch $b --steps 2 --warmup 3 --no-cpu 2>/dev/null | tail -1 >> gpurun_out/batch_sweep_$tag.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/batch_sweep_$tag.jsonl"):
    d = json.loads(l)
    print("batch", d["config"]["batch_per_gpu"], "samples/s %.2f" % d["value"], "e2e %.2f" % d["e2e"]["value"], "step ms %.3f" % d["unet_step_ms"],
          "conv TFLOP/s %.0f" % d["roofline"]["achieved"], "frac %.3f" % d["roofline"]["frac"])
PY
