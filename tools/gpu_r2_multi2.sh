#!/bin/bash
# Round 2, later kernels: training and U-Net lines at N GPUs (tools/gpu_r2_multi2.sh N under `gpurun --gpus N`)
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --workload train --no-cpu-baseline > gpurun_out/r2_train_${N}gpu.json 2> gpurun_out/r2_multi2_${N}.err; echo "train rc=$?"
run --workload unet1d --no-cpu-baseline > gpurun_out/r2_unet_${N}gpu.json 2>> gpurun_out/r2_multi2_${N}.err; echo "unet rc=$?"
tail -3 gpurun_out/r2_multi2_${N}.err
python - <<PY
import json
for f in ("r2_train_${N}gpu", "r2_unet_${N}gpu"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"], 1), d["unit"], "frac", round(d["roofline"]["frac"], 3), "per_rank" in d and [round(r["ms_per_step"], 1) for r in d["per_rank"]])
PY
