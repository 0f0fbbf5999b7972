#!/bin/bash
# quick loop: fused GroupNorm conv parity + U-Net timing fused vs unfused
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gn_conv.py tests/test_gpu_unet1d.py -q -x 2>&1 | tail -4
{
timeout 300 python tools/time_unet.py 128 262144 bf16 10
ADB_UNET_NOFUSE=1 timeout 300 python tools/time_unet.py 128 262144 bf16 10
timeout 300 python tools/time_unet.py 32 262144 bf16 10
} > gpurun_out/r2s_time.log 2>&1; cat gpurun_out/r2s_time.log
if [ "$1" = "ncu" ]; then
ADB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_unet1d_b128.csv python tools/time_unet.py 128 262144 bf16 1 > gpurun_out/r2s_ncu.log 2>&1
echo "ncu rc=$?"
fi
