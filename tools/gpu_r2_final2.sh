#!/bin/bash
# Round 2 evidence refresh on one GPU (later kernels): whole GPU suite, headline bench, reference arm, training and U-Net lines.
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_final2.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_final2.err; echo "ref rc=$?"
timeout 600 python bench.py --workload train > gpurun_out/r2_train_1gpu.json 2>> gpurun_out/r2_final2.err; echo "train rc=$?"
timeout 900 python bench.py --workload unet1d > gpurun_out/r2_unet_1gpu.json 2>> gpurun_out/r2_final2.err; echo "unet rc=$?"
tail -3 gpurun_out/r2_final2.err
python - <<PY
import json
for f in ("r2_bench_1gpu", "r2_bench_reference_arm", "r2_train_1gpu", "r2_unet_1gpu"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 3), d["unit"], d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
