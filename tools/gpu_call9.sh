#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r01_v24}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q > gpurun_out/${TAG}_pytest_render.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_render.log
tail -6 gpurun_out/${TAG}_pytest_render.log
timeout 200 python tools/gpu_render_ab.py 600 > gpurun_out/${TAG}_render_ab.txt 2>&1
tail -4 gpurun_out/${TAG}_render_ab.txt
