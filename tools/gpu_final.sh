#!/bin/bash
# Round-end evidence run: tests, the three bench lines, ncu launch list of the headline bench and a full capture of its top kernel.
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ref rc=$?"
python bench.py --workload train > gpurun_out/bench_train_$tag.json 2>> gpurun_out/bench_$tag.err; echo "train rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4700 -c 1600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
python tools/time_net.py 256 2 1 > gpurun_out/plain_tn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wavenet_block -s 2 -c 2 -o gpurun_out/prof_block_$tag python tools/time_net.py 256 2 1 > gpurun_out/ncu_tn.log 2>&1
echo "ncu full rc=$?"
python tools/time_step.py > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none -k regex:edm_kernel -s 4 -c 2 -o gpurun_out/prof_step_$tag python tools/time_step.py 67108864 2 > gpurun_out/ncu_step.log 2>&1
echo "ncu step rc=$?"; cat gpurun_out/plain_step.log
python bench.py --workload unet1d > gpurun_out/bench_unet_$tag.json 2>> gpurun_out/bench_$tag.err; echo "unet rc=$?"
