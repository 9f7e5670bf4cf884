#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, the ncu launch list of a short bench
# run, and --set full captures of the two hot kernels.  usage: tools/gpu_validate.sh <tag> [nocap]
cd "$(dirname "$0")/.."
TAG=${1:-r01_v8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-render --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_under_ncu.json 2>&1
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.summary.txt 2>&1
tail -30 gpurun_out/${TAG}_launches.summary.txt
if [ "$2" != "nocap" ]; then bash tools/ncu_v8.sh ${TAG}; fi
rm -f gpurun_out/${TAG}_launches.csv.gz; gzip -f gpurun_out/${TAG}_launches.csv
