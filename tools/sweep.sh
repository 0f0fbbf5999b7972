#!/bin/bash
# BASELINE.json configs[4]: batch sweep of the sampling path on this box's GPU(s). Usage: tools/sweep.sh <tag>
tag=${1:-r1}
out=gpurun_out/sweep_$tag.jsonl
: > $out
for b in 64 128 256 512 1024; do
  python bench.py --batch $b --steps 1 --warmup 3 --no-cpu-baseline >> $out 2>> gpurun_out/sweep_$tag.err
done
python - <<PY
import json
for line in open("$out"):
    d = json.loads(line)
    print(d["config"]["batch_per_gpu"], round(d["value"], 2), "samples/s", round(d["roofline"]["achieved"], 1), "TFLOP/s conv", d["clocks"]["sm_mhz"], "MHz")
PY
