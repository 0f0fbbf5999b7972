#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gn_conv.py -q -x 2>&1 | tail -4
{ timeout 300 python tools/time_gnconv.py 128; ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_gnconv.py 128 | grep -A14 "res=False" | head -34; } > gpurun_out/r2u_gnconv.log 2>&1
cat gpurun_out/r2u_gnconv.log
timeout 300 python tools/time_unet.py 128 262144 bf16 10; timeout 600 python -m pytest tests/test_gpu_unet1d.py -q -x 2>&1 | tail -3
