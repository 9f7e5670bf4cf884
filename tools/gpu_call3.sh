#!/bin/bash
# microbench of the Ozaki GEMM, GPU parity tests, short bench (no render / CPU baseline)
cd "$(dirname "$0")/.."
TAG=${1:-r01_v10}
mkdir -p gpurun_out
timeout 300 tools/microbench/bin/ozaki_test > gpurun_out/${TAG}_ozaki_microbench.txt 2>&1
grep "dbg=\|PASS\|FAIL" gpurun_out/${TAG}_ozaki_microbench.txt | cut -c1-60,150-260
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --no-render --no-cpu-baseline > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench_short.json; tail -5 gpurun_out/${TAG}_bench.err
