#!/bin/bash
# Round 2, session L: skip GEMM with the fused tail: parity (all wavenet tests) + timing A/B + headline bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wavenet.py tests/test_gpu_abi_errors.py -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2l_pytest.log
{
timeout 300 python tools/time_net.py 256 36 5
ADB_FUSE_TAIL=0 timeout 300 python tools/time_net.py 256 36 5
} > gpurun_out/r2l_time.log 2>&1; cat gpurun_out/r2l_time.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2l_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2l_bench.json"))
print(round(d["value"], 2), "samples/s", d["roofline"]["frac"], d["roofline"]["residual_stack"]["frac"], d["kernel_ms"], d["clocks"])
PY
