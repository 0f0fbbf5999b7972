#!/bin/bash
# One gpurun call: parity tests, cycle accounting, bench. Usage: tools/gpu_session.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$tag.log
{
  for pair in 0 1; do
    ADB_TC_PAIR=$pair timeout 300 python tools/time_net.py 64 36 3
    ADB_TC_PAIR=$pair ADB_DEBUG_FLAGS=2 timeout 300 python tools/time_net.py 64 36 3
  done
  ADB_TC_PAIR=0 ADB_TC_CLUSTER=1 timeout 300 python tools/time_net.py 64 36 3
} > gpurun_out/timenet_$tag.log 2>&1
cat gpurun_out/timenet_$tag.log
