#!/bin/bash
mkdir -p gpurun_out
{
echo "=== accounting ML=2 PIPE=0"
ADB_ZS_ML=2 ADB_ZS_PIPE=0 ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 256 36 2 2>&1 | tail -14
echo "=== accounting ML=0 PIPE=0"
ADB_ZS_ML=0 ADB_ZS_PIPE=0 ADB_LIB=debug ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 256 36 2 2>&1 | tail -14
echo "=== dram bytes, ML=2 (one launch = 36 blocks)"
ADB_ZS_ML=2 ADB_ZS_PIPE=0 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:wavenet_block_zs -s 1 -c 1 python tools/time_net.py 256 36 1 2>&1 | grep -E "dram__|gpu__time|lts__|wavenet_block"
echo "=== dram bytes, ML=0, one block launch"
ADB_ZS_ML=0 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:wavenet_block_zs -s 40 -c 1 python tools/time_net.py 256 36 1 2>&1 | grep -E "dram__|gpu__time|lts__|wavenet_block"
} > gpurun_out/r2ac.log 2>&1
cat gpurun_out/r2ac.log
