#!/bin/bash
# Jacobi warm-start gate (EMAGLS_JACOBI_WARM_GRADING): arbitration table (err_cuda against exact arithmetic), design
# parity tests and a short bench line per value.  usage: tools/gpu_r2_warm.sh <tag> <value> [<value> ...]
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  export EMAGLS_JACOBI_WARM_GRADING=$v
  timeout 600 python -m pytest tests/test_gpu_arbitration.py tests/test_gpu_design.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest_$v.log 2>&1
  echo "== grading $v: pytest exit $?"; tail -2 gpurun_out/${TAG}_pytest_$v.log
  grep -E "^c1:" gpurun_out/${TAG}_pytest_$v.log | head -16
  grep -E "^c5:" gpurun_out/${TAG}_pytest_$v.log | head -8
  timeout 300 python bench.py --steps 3 --warmup 3 --no-render --no-cpu-baseline --no-spot-check > gpurun_out/${TAG}_bench_$v.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_bench_$v.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("grading $v:", round(j["value"], 1), round(j["ms_per_step"], 1), "jacobi", round(r["classes"]["jacobi"]["ms_per_step"], 1), "sweeps", round(r["jacobi_mean_sweeps"], 2))
PY
done
