#!/bin/bash
# N-GPU run of the bench contract under torchrun (weak-scaling line + strong_scaling + sharded render e2e).
# usage: tools/gpu_ngpu.sh <tag> <N>
cd "$(dirname "$0")/.."
TAG=${1:-r02}; N=${2:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/${TAG}_${N}gpu_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-spot-check --render-seconds 120 \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "exit $?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("N", j["n_gpus"], "value", round(j["value"], 1), "ms", round(j["ms_per_step"], 1), "e2e", round(j["e2e"]["value"], 1),
          "gather", j["e2e"].get("includes_nccl_gather"))
    print("strong", j.get("strong_scaling"))
    print("render e2e", (j.get("render") or {}).get("e2e"))
    print("clocks", j.get("clocks"))
except Exception as e:
    print("parse failed", repr(e)); print(open("gpurun_out/${TAG}_bench_${N}gpu.err").read()[-2500:])
PY
