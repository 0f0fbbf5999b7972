#!/bin/bash
# Training step: bench line + ncu launch list of the steps at B=32.
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2y_train.json 2> gpurun_out/r2y_train.err; echo "train rc=$?"; tail -2 gpurun_out/r2y_train.err
python -c "
import json; d=json.load(open('gpurun_out/r2y_train.json')); print(round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms/step', round(d['roofline']['frac'],3), d['gpu_launches'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python bench.py --workload train --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r2y_ncu.log 2>&1
echo "ncu rc=$?"
