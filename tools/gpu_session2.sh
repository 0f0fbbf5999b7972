#!/bin/bash
tag=${1:-r1}
mkdir -p gpurun_out
{
  for f in 0 4 6; do ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_net.py 64 36 3; done
  for f in 0 4; do ADB_DEBUG_FLAGS=$f timeout 300 python tools/time_net.py 256 36 2; done
} > gpurun_out/timenet_$tag.log 2>&1
cat gpurun_out/timenet_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4700 -c 1600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
python tools/time_net.py 64 4 1 > gpurun_out/plain_tn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wavenet_block -s 4 -c 3 -o gpurun_out/prof_block_$tag python tools/time_net.py 64 4 1 > gpurun_out/ncu_tn.log 2>&1
echo "ncu full rc=$?"
