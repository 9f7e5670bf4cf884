#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_design.py tests/test_gpu_variants.py tests/test_gpu_blocks.py -m gpu -x -q > gpurun_out/r01_v27_pytest_design.log 2>&1
echo "pytest exit $?" >> gpurun_out/r01_v27_pytest_design.log
tail -6 gpurun_out/r01_v27_pytest_design.log
