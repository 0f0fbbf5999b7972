#!/bin/bash
# Round 2, session I: fp16 x bf16 weight-gradient GEMM (no re-typing of the z stash): parity, then bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r2i_pytest_mixed.log 2>&1; echo "pytest mixed rc=$?"; tail -4 gpurun_out/r2i_pytest_mixed.log
ADB_WGRAD_MIXED=0 timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r2i_pytest_cast.log 2>&1; echo "pytest cast rc=$?"; tail -2 gpurun_out/r2i_pytest_cast.log
timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2i_train_mixed.json 2> gpurun_out/r2i_train.err; echo "train mixed rc=$?"
ADB_WGRAD_MIXED=0 timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2i_train_cast.json 2>> gpurun_out/r2i_train.err; echo "train cast rc=$?"
python - <<PY
import json
for f in ("r2i_train_mixed", "r2i_train_cast"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 2), "ms/step", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
