#!/bin/bash
mkdir -p gpurun_out
{ timeout 300 python tools/time_gnconv.py 128; ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_gnconv.py 128; } > gpurun_out/r2t_gnconv.log 2>&1
cat gpurun_out/r2t_gnconv.log
