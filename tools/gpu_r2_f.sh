#!/bin/bash
# Round 2, session F: GPU suite after the Philox / generic-grad / EMA / U-Net-trajectory changes + U-Net bench line.
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --workload unet1d --no-cpu-baseline > gpurun_out/r2f_bench_unet.json 2> gpurun_out/r2f_bench_unet.err; echo "unet bench rc=$?"; tail -3 gpurun_out/r2f_bench_unet.err
cut -c1-1500 gpurun_out/r2f_bench_unet.json
