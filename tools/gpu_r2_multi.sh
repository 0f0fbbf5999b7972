#!/bin/bash
# Round 2 multi-GPU records: tools/gpu_r2_multi.sh N  (run under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
if [ "$N" = "1" ]; then run() { python bench.py "$@"; }; fi
run --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_multi_${N}.err; echo "headline rc=$?"
run --workload sweep --steps 1 --no-cpu-baseline > gpurun_out/r2_sweep_${N}gpu.jsonl 2>> gpurun_out/r2_multi_${N}.err; echo "sweep rc=$?"
run --workload train --no-cpu-baseline > gpurun_out/r2_train_${N}gpu.json 2>> gpurun_out/r2_multi_${N}.err; echo "train rc=$?"
if [ "$N" = "1" ] || [ "$N" = "8" ]; then
  run --workload unet1d --no-cpu-baseline > gpurun_out/r2_unet_${N}gpu.json 2>> gpurun_out/r2_multi_${N}.err; echo "unet rc=$?"
fi
tail -3 gpurun_out/r2_multi_${N}.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_*_${N}gpu.json*")):
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(f.split("/")[-1], d["config"].get("global_batch"), round(d["value"], 2), d["unit"], "frac", round(d["roofline"]["frac"], 3),
              "per_rank" in d and [round(r["ms_per_step"], 1) for r in d["per_rank"]])
PY
