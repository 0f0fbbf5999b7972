#!/bin/bash
# Round 2, session E: whole GPU test suite + headline bench with the z-stash path.
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2e_pytest.log
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err
cat gpurun_out/r2e_bench.json | cut -c1-3000
