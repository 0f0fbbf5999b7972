#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_gn_conv.py -q -x 2>&1 | tail -2
for f in 2 18; do echo "=== flags $f"; ADB_LIB=debug ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_gnconv.py 128 2>&1 | grep -A14 "Cin=256+0 N=256 res=False" | grep -v "epi:t" | head -14; done
timeout 300 python tools/time_gnconv.py 128
timeout 300 python tools/time_unet.py 128 262144 bf16 10
echo "=== z-stash kernel warp roles"
ADB_ZS_HI_ROLES=0 timeout 300 python tools/time_net.py 256 36 3
ADB_ZS_HI_ROLES=1 timeout 300 python tools/time_net.py 256 36 3
ADB_ZS_HI_ROLES=0 timeout 300 python tools/time_net.py 256 36 3
ADB_ZS_HI_ROLES=1 timeout 300 python tools/time_net.py 256 36 3
ADB_ZS_HI_ROLES=1 timeout 300 python -m pytest tests/test_gpu_wavenet.py -q -x 2>&1 | tail -2
} > gpurun_out/r2w.log 2>&1
cat gpurun_out/r2w.log
