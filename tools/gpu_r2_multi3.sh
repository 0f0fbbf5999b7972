#!/bin/bash
# training line at N GPUs with the last build (tools/gpu_r2_multi3.sh N under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train --no-cpu-baseline > gpurun_out/r2_train_${N}gpu.json 2> gpurun_out/r2_multi3_${N}.err; echo "train rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_train_${N}gpu.json')); print(d['n_gpus'], round(d['value'],1), d['unit'], 'frac', round(d['roofline']['frac'],3))"
