#!/bin/bash
# Round 2, session M: one-launch refold + faster rank-B correction: all wavenet / training / sampler parity tests, training bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_wavenet.py tests/test_extra_samplers_ema.py tests/test_dpm_unipc.py tests/test_wav_module.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.log
timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2m_train.json 2> gpurun_out/r2m_train.err; echo "train rc=$?"; tail -2 gpurun_out/r2m_train.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2m_train.json"))
print(round(d["value"], 1), "samples/s", round(d["ms_per_step"], 2), "ms/step", round(d["roofline"]["frac"], 3), d["gpu_launches"], "launches")
PY
