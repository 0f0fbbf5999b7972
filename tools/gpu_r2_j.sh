#!/bin/bash
# Round 2, session J: weight gradient on h with the rank-B step-embedding correction (no add_bcast pass): parity + bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2j_train.json 2> gpurun_out/r2j_train.err; echo "train rc=$?"; tail -2 gpurun_out/r2j_train.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2j_train.json"))
print(round(d["value"], 1), "samples/s", round(d["ms_per_step"], 2), "ms/step", round(d["roofline"]["frac"], 3), d["gpu_launches"], "launches")
PY
