#!/bin/bash
# Premise check 2: the round-1 pair kernel with the fp32 skip read-modify-write in every block (ADB_BLOCK_KERNEL=1), all
# activation / skip accesses confined to a window of W samples (flag 2048, W = flags >> 16) = the traffic pattern of a
# multi-layer kernel whose pass stays in L2.
mkdir -p gpurun_out
{
for f in 0 $((2048 + (1<<16))) $((2048 + (2<<16))) $((2048 + (3<<16))); do
  ADB_BLOCK_KERNEL=1 ADB_LIB=debug ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_net.py 256 36 3
done
ADB_BLOCK_KERNEL=2 ADB_LIB=debug ADB_DEBUG_FLAGS=0 timeout 300 python tools/time_net.py 256 36 3
ADB_LIB=debug ADB_DEBUG_FLAGS=0 timeout 300 python tools/time_net.py 256 36 3
} > gpurun_out/r2p_time.log 2>&1; cat gpurun_out/r2p_time.log
