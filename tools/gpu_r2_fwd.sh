#!/bin/bash
# Forward-product epilogue A/B: microbenchmark of the three epilogues (digit checksums must agree), the parity tests
# that go through the recursion, and short bench lines.  usage: tools/gpu_r2_fwd.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-r02_fwd}
mkdir -p gpurun_out
{
for mode in scaled raw tma; do
  if [ $mode = tma ]; then unset EMAGLS_OZ_FWD; else export EMAGLS_OZ_FWD=$mode; fi
  timeout 120 tools/microbench/bin/oz_fwd_bench 3600 1 20 6
  timeout 120 tools/microbench/bin/oz_fwd_bench 900 3 20 6
  timeout 120 tools/microbench/bin/oz_fwd_bench 450 1 20 4
done
unset EMAGLS_OZ_FWD
} > gpurun_out/${TAG}_fwd_microbench.txt 2>&1
cat gpurun_out/${TAG}_fwd_microbench.txt
timeout 900 python -m pytest tests/test_gpu_design.py tests/test_gpu_variants.py tests/test_gpu_arbitration.py -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
for kv in A=1 EMAGLS_OZ_FWD=raw; do
  env $kv timeout 600 python bench.py --steps 3 --warmup 3 --no-render --no-cpu-baseline --no-spot-check > gpurun_out/${TAG}_bench_${kv//[^A-Za-z0-9_]/_}.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_${kv//[^A-Za-z0-9_]/_}.json").read().strip().splitlines()[-1])
    print("AB ${kv}", round(j["value"], 1), round(j["ms_per_step"], 1), {k: round(v["ms_per_step"], 1) for k, v in j["roofline"]["classes"].items()})
except Exception as e:
    print("ab parse failed", repr(e)); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
done
