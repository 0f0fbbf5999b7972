#!/bin/bash
# Round 2, session C: what makes the TMA stores expensive? (timing-only variants of the debug build at B=256)
mkdir -p gpurun_out
{
for f in 0 1024 256 1280 32; do
  ADB_LIB=debug ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_net.py 256 36 5
done
ADB_LIB=debug ADB_DEBUG_FLAGS=34 timeout 300 python tools/time_net.py 256 36 2
ADB_LIB=debug ADB_DEBUG_FLAGS=26 timeout 300 python tools/time_net.py 256 36 2
} > gpurun_out/r2c_time.log 2>&1
cat gpurun_out/r2c_time.log
