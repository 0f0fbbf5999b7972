#!/bin/bash
# Fused GroupNorm-apply convolution: op-level parity with both descriptor variants, then the U-Net goldens.
mkdir -p gpurun_out
for boff in 0 1; do
  echo "=== ADB_GC_BOFF=$boff"
  ADB_GC_BOFF=$boff timeout 600 python -m pytest tests/test_gpu_gn_conv.py -q -s 2>&1 | grep -v "^$" | tail -25
done > gpurun_out/r2q_gnconv.log 2>&1
cat gpurun_out/r2q_gnconv.log
timeout 900 python -m pytest tests/test_gpu_unet1d.py -q -x 2>&1 | tail -5
