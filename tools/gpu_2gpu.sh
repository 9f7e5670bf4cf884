#!/bin/bash
# 2-GPU run of the bench contract (torchrun, NCCL gather of the banks) + the reference arm under torchrun
cd "$(dirname "$0")/.."
TAG=${1:-r01_v15}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_2gpu_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --render-seconds 60 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
echo "exit $?"; tail -c 2500 gpurun_out/${TAG}_bench_2gpu.json; tail -8 gpurun_out/${TAG}_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref_2gpu.json 2>> gpurun_out/${TAG}_bench_2gpu.err
echo "exit $?"; tail -c 600 gpurun_out/${TAG}_bench_ref_2gpu.json
