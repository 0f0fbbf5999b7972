#!/bin/bash
# U-Net with the fused GroupNorm-apply convolution: goldens, timing fused vs unfused, launch list of one evaluation.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet1d.py tests/test_gpu_gn_conv.py -q -x 2>&1 | tail -5
{
timeout 300 python tools/time_unet.py 128 262144 bf16 10
ADB_UNET_NOFUSE=1 timeout 300 python tools/time_unet.py 128 262144 bf16 10
timeout 300 python tools/time_unet.py 32 262144 bf16 10
ADB_UNET_NOFUSE=1 timeout 300 python tools/time_unet.py 32 262144 bf16 10
timeout 300 python tools/time_unet.py 128 262144 bf16 10
} > gpurun_out/r2r_time.log 2>&1; cat gpurun_out/r2r_time.log
ADB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_unet1d_b128.csv python tools/time_unet.py 128 262144 bf16 1 > gpurun_out/r2r_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2r_ncu.log
