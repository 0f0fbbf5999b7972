#!/bin/bash
# Round 2 evidence (later kernels): ncu launch lists of a training step and of one U-Net evaluation, full captures of the fused
# GroupNorm convolution, the pair convolution and the pair weight-gradient kernel.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_bf16_b32.csv python bench.py --workload train --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu2_train.log 2>&1; echo "ncu train rc=$?"
ADB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_unet1d_b128.csv python tools/time_unet.py 128 262144 bf16 1 > gpurun_out/ncu2_unet.log 2>&1; echo "ncu unet rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cl_conv3_gn_tc|cl_conv_tc_kernel" -s 4 -c 4 -o gpurun_out/r2_prof_gnconv python tools/time_gnconv.py 128 > gpurun_out/ncu2_gnconv.log 2>&1; echo "ncu gnconv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_pair" -s 40 -c 2 -o gpurun_out/r2_prof_wgrad python bench.py --workload train --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu2_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
