#!/bin/bash
# Round 2, session G (1 GPU): GPU suite, batch sweep 64..4096 on one GPU (configs[4]), reference classes on the GPU.
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
timeout 1500 python bench.py --workload sweep --steps 1 --no-cpu-baseline > gpurun_out/r2_sweep_1gpu.jsonl 2> gpurun_out/r2g_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for line in open("gpurun_out/r2_sweep_1gpu.jsonl"):
    d = json.loads(line)
    print(d["config"]["global_batch"], round(d["value"], 2), "samples/s", round(d["roofline"]["achieved"], 1), "TFLOP/s block", d["clocks"]["sm_mhz"], "MHz")
PY
timeout 900 python tools/ref_gpu.py 64 256 > gpurun_out/r2_reference_gpu.jsonl 2> gpurun_out/r2g_ref.err; echo "ref gpu rc=$?"; cat gpurun_out/r2_reference_gpu.jsonl; tail -3 gpurun_out/r2g_ref.err
