#!/bin/bash
# Round 2, session H: training step with the forward keeping y / z (no recompute in the backward): parity + bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_wavenet.py -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2h_train_zs.json 2> gpurun_out/r2h_train.err; echo "train zs rc=$?"
ADB_BLOCK_KERNEL=2 timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2h_train_pair.json 2>> gpurun_out/r2h_train.err; echo "train pair rc=$?"
tail -3 gpurun_out/r2h_train.err
python - <<PY
import json
for f in ("r2h_train_zs", "r2h_train_pair"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 2), "ms/step", round(d["roofline"]["frac"], 3), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
