#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gn_conv.py tests/test_gpu_unet1d.py tests/test_gpu_sharding_2proc.py -q -x -s 2>&1 | grep -v "^$" | tail -22
{ timeout 300 python tools/time_gnconv.py 128; ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_gnconv.py 128 | grep -A14 "Cin=256+0 N=256 res=False" | head -16; } 2>&1 | tee gpurun_out/r2aa_gnconv.log
timeout 300 python tools/time_unet.py 128 262144 bf16 10
