#!/bin/bash
# CTA-pair weight-gradient kernel + CTA-pair generic convolution: whole GPU suite, training bench, U-Net timing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2z_train.json 2> gpurun_out/r2z_train.err; echo "train rc=$?"; tail -2 gpurun_out/r2z_train.err
python -c "
import json; d=json.load(open('gpurun_out/r2z_train.json')); print(round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms/step', round(d['roofline']['frac'],3), d['gpu_launches'])"
timeout 300 python tools/time_unet.py 128 262144 bf16 10
timeout 300 python tools/time_gnconv.py 128 | head -3
