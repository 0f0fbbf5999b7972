#!/bin/bash
# launch lists of the last build (training step at B=32, three U-Net evaluations at B=128)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_bf16_b32.csv python bench.py --workload train --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu3_train.log 2>&1; echo "ncu train rc=$?"
ADB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_unet1d_b128.csv python tools/time_unet.py 128 262144 bf16 1 > gpurun_out/ncu3_unet.log 2>&1; echo "ncu unet rc=$?"
