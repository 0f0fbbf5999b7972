#!/bin/bash
# Round 2 evidence: ncu launch list of the headline bench and full captures of the z-stash block kernel and the skip GEMM.
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
python tools/time_net.py 256 4 1 > gpurun_out/plain_tn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wavenet_block_zs|wavenet_skip_gemm" -s 5 -c 5 -o gpurun_out/r2_prof_zs python tools/time_net.py 256 4 1 > gpurun_out/ncu_tn.log 2>&1
echo "ncu full rc=$?"; cat gpurun_out/plain_tn.log
