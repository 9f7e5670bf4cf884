#!/bin/bash
# 4-GPU run of the bench contract (weak scaling: 3600 orientations per rank, NCCL gather in e2e)
cd "$(dirname "$0")/.."
TAG=${1:-r01_v23}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/${TAG}_bench_4gpu.json 2> gpurun_out/${TAG}_bench_4gpu.err
echo "exit $?"; tail -c 1500 gpurun_out/${TAG}_bench_4gpu.json | cut -c1-900; grep -i "error\|Traceback" gpurun_out/${TAG}_bench_4gpu.err | head -5
