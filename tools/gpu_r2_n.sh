#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/time_net.py 256 36 5
ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 256 36 2
timeout 300 python tools/time_net.py 256 36 5
} > gpurun_out/r2n_time.log 2>&1; cat gpurun_out/r2n_time.log
timeout 300 python -m pytest tests/test_gpu_wavenet.py -x -q 2>&1 | tail -2
