#!/bin/bash
# second half of the final validation: full launch list of one step + captures of the Gram-route kernels
cd "$(dirname "$0")/.."
TAG=${1:-r01_v18}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 3700 \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-render --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_under_ncu.json 2>&1
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.summary.txt 2>&1
head -24 gpurun_out/${TAG}_launches.summary.txt
rm -f gpurun_out/${TAG}_launches.csv.gz; gzip -f gpurun_out/${TAG}_launches.csv
cap() { # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
      python tools/gpu_ncu_factor.py 3600 > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
cap ${TAG}_gram_sweep "gram_sweep_kernel" 1 1
cap ${TAG}_bwd_small "bwd_small_kernel" 60 1
grep -n "Duration\|Executed Ipc Active\|Issue Slots Busy\|DRAM Throughput\|Achieved Occupancy\|Registers Per" gpurun_out/${TAG}_gram_sweep.details.txt gpurun_out/${TAG}_bwd_small.details.txt
