#!/bin/bash
# Round 2, session A: parity of the z-stash path, then block-kernel timing (z-stash vs pair) with cycle accounting.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wavenet.py -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
{
for k in 3 2; do
  ADB_BLOCK_KERNEL=$k timeout 300 python tools/time_net.py 64 36 3
done
ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 64 36 3
for k in 3 2; do
  ADB_BLOCK_KERNEL=$k timeout 300 python tools/time_net.py 256 36 2
done
ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 256 36 2
} > gpurun_out/r2a_time.log 2>&1
cat gpurun_out/r2a_time.log
