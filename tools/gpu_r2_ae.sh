#!/bin/bash
# programmatic dependent launch on the U-Net's hot kernels: parity, then A/B timing
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_gn_conv.py tests/test_gpu_unet1d.py tests/test_gpu_train.py -q -x 2>&1 | tail -3
for pdl in 1 0 1 0; do echo "ADB_PDL=$pdl"; ADB_PDL=$pdl timeout 300 python tools/time_unet.py 128 262144 bf16 10; done
for pdl in 1 0; do echo "ADB_PDL=$pdl"; ADB_PDL=$pdl timeout 300 python tools/time_unet.py 32 262144 bf16 10; done
ADB_PDL=1 timeout 300 python bench.py --workload train --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train pdl=1', round(d['value'],1))"
ADB_PDL=0 timeout 300 python bench.py --workload train --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train pdl=0', round(d['value'],1))"
} > gpurun_out/r2ae.log 2>&1
cat gpurun_out/r2ae.log
