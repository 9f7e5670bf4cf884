#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r01_v23}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -1
