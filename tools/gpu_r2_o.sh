#!/bin/bash
# Premise check for a multi-layer L2-resident block kernel: timing-only variants of the z-stash kernel whose activation
# loads / residual reads / h' stores are confined to a window of W samples (flag 2048, W = flags >> 16).
mkdir -p gpurun_out
{
for f in 0 $((2048 + (2<<16))) $((2048 + (4<<16))) $((2048 + (8<<16))) $((2048 + 8 + (4<<16))) 0; do
  ADB_LIB=debug ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_net.py 256 36 3
done
} > gpurun_out/r2o_time.log 2>&1; cat gpurun_out/r2o_time.log
