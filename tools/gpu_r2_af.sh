#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wavenet.py -q -x -k "multilayer" 2>&1 | tail -2
ADB_ZS_ML=4 timeout 600 python -m pytest tests/test_gpu_wavenet.py -q -x -k "full_size_batch or batch_rows" 2>&1 | tail -2
for cfg in "0 1" "4 1" "5 1" "4 0" "0 1" "4 1"; do
  set -- $cfg
  echo "=== ADB_ZS_ML=$1 ADB_ZS_PIPE=$2"
  ADB_ZS_ML=$1 ADB_ZS_PIPE=$2 timeout 300 python tools/time_net.py 256 36 3 2>&1 | tail -1 | cut -c1-150
done
} > gpurun_out/r2af.log 2>&1
cat gpurun_out/r2af.log
