#!/bin/bash
# Round 2, session K: launch list of the training step (where the backward's time goes now) + 1-GPU sweep rerun.
mkdir -p gpurun_out
python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file gpurun_out/r2_launches_train.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
echo "ncu train rc=$?"
timeout 1500 python bench.py --workload sweep --steps 1 --no-cpu-baseline > gpurun_out/r2_sweep_1gpu.jsonl 2> gpurun_out/r2k_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for line in open("gpurun_out/r2_sweep_1gpu.jsonl"):
    d = json.loads(line)
    print(d["config"]["global_batch"], round(d["value"], 2), "samples/s", round(d["roofline"]["achieved"], 1), "TFLOP/s block", d["clocks"]["sm_mhz"], "MHz")
PY
