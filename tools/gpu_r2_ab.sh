#!/bin/bash
# Multi-layer wavefront launch of the z-stash block kernel (ADB_ZS_ML = samples per L2-resident sub-pass): parity, then timing.
mkdir -p gpurun_out
{
ADB_ZS_ML=2 timeout 600 python -m pytest tests/test_gpu_wavenet.py -q -x 2>&1 | tail -4
for cfg in "0 1" "2 0" "3 0" "4 0" "2 1" "4 1" "0 1"; do
  set -- $cfg
  echo "=== ADB_ZS_ML=$1 ADB_ZS_PIPE=$2"
  ADB_ZS_ML=$1 ADB_ZS_PIPE=$2 timeout 300 python tools/time_net.py 256 36 3 2>&1 | tail -1
done
} > gpurun_out/r2ab.log 2>&1
cat gpurun_out/r2ab.log
