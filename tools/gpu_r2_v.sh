#!/bin/bash
mkdir -p gpurun_out
for f in 2 6 10 18; do echo "=== flags $f"; ADB_LIB=debug ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_gnconv.py 128 2>&1 | grep -A14 "Cin=256+0 N=256 res=False" | grep -v "prod:\|epi:w\|epi:t" | head -12; done > gpurun_out/r2v.log 2>&1
cat gpurun_out/r2v.log
