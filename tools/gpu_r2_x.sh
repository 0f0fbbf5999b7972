#!/bin/bash
# U-Net evidence after the fused GroupNorm convolution: whole GPU suite, U-Net bench line, launch list of one evaluation.
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --workload unet1d > gpurun_out/r2_unet_1gpu.json 2> gpurun_out/r2x.err; echo "unet rc=$?"; tail -2 gpurun_out/r2x.err
python -c "
import json; d=json.load(open('gpurun_out/r2_unet_1gpu.json')); print(d['value'], d['ms_per_network_evaluation'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
ADB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_unet1d_b128.csv python tools/time_unet.py 128 262144 bf16 1 > gpurun_out/r2x_ncu.log 2>&1
echo "ncu rc=$?"
