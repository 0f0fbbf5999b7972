#!/bin/bash
# Round 2, session B: where does the z-stash kernel's time / energy go? Timing-only variants of the debug build at B=256.
mkdir -p gpurun_out
{
for f in 0 4 8 16 24 32 64 124; do
  ADB_LIB=debug ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_net.py 256 36 5
done
ADB_BLOCK_KERNEL=2 timeout 300 python tools/time_net.py 256 36 5
timeout 300 python tools/time_net.py 256 36 5
} > gpurun_out/r2b_time.log 2>&1
cat gpurun_out/r2b_time.log
timeout 600 python -m pytest tests/test_gpu_wavenet.py -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
