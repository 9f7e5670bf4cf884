#!/bin/bash
# GPU call 2 of the session: microbench A/B of the Ozaki GEMM (8 vs 16 epilogue warps, tile order),
# GPU parity tests incl. the new front-end rows, short bench A/B, default bench.
cd "$(dirname "$0")/.."
TAG=${1:-r01_v9}
mkdir -p gpurun_out
( echo "== 16 epilogue warps, n-fastest tile order for the backward shape"; timeout 300 tools/microbench/bin/ozaki_test;
  echo "== 8 epilogue warps"; timeout 300 tools/microbench/bin/ozaki_test_w8 ) > gpurun_out/${TAG}_ozaki_microbench.txt 2>&1
tail -32 gpurun_out/${TAG}_ozaki_microbench.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
EMAGLS_LIB_PATH=$PWD/tools/microbench/bin/libemagls_cuda_w8.so timeout 600 python bench.py --no-render --no-cpu-baseline > gpurun_out/${TAG}_bench_w8.json 2> gpurun_out/${TAG}_bench.err
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
tail -c 2500 gpurun_out/${TAG}_bench_w8.json; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
