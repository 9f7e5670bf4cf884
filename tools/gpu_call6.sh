#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r01_v17}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
for v in 64 96; do
EMAGLS_FACTOR_RB=$v timeout 600 python bench.py --no-render --no-cpu-baseline > gpurun_out/${TAG}_bench_rb$v.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_rb$v.json"))
r=d["roofline"]
print("RB=$v", d["value"], d["e2e"]["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"])
print({k: round(v*d["ms_per_step"],1) for k,v in r["class_time_share"].items()})
PY
done
tail -5 gpurun_out/${TAG}_bench.err
