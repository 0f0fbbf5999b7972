#!/bin/bash
# Round 2, session D: pipelined job order of the z-stash kernel — parity, then timing against the plain order.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wavenet.py -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
ADB_ZS_PIPE=0 timeout 600 python -m pytest tests/test_gpu_wavenet.py -x -q -k "block_kernels or zstash" > gpurun_out/r2d_pytest_nopipe.log 2>&1; echo "pytest nopipe rc=$?"; tail -3 gpurun_out/r2d_pytest_nopipe.log
{
for pz in 1 0; do
  ADB_ZS_PIPE=$pz timeout 300 python tools/time_net.py 256 36 5
done
ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 256 36 2
ADB_ZS_PIPE=0 ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 256 36 2
timeout 300 python tools/time_net.py 64 36 5
} > gpurun_out/r2d_time.log 2>&1
cat gpurun_out/r2d_time.log
