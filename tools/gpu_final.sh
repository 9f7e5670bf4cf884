#!/bin/bash
# Final validation of a tree: GPU parity tests, default bench line, reference arm, ncu launch list of one
# step, --set full captures of the hot kernels.  usage: tools/gpu_final.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-r01_v18}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
cap() { # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
      python tools/gpu_ncu_factor.py 3600 > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
cap ${TAG}_oz "ozaki_gemm_kernel" 40 2
cap ${TAG}_factor "factor_kernel" 2 1
cap ${TAG}_gram "gram_sweep_kernel\|bwd_small_kernel" 2 2
python tools/ncu_traffic.py gpurun_out/${TAG}_oz.raw.csv profiles/r01_oz_traffic.json "ncu --set full --clock-control none, ozaki_gemm_kernel<6> launches 41-42 of a 3600-orientation design (forward EpiPhaseSlice 128x64 tiles, then backward EpiStoreF64 128x80 tiles); tools/gpu_final.sh ${TAG}" > /dev/null 2>&1
cp profiles/r01_oz_traffic.json gpurun_out/${TAG}_oz_traffic.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3200 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
tail -c 700 gpurun_out/${TAG}_bench_ref.json
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
# launch list: the warm-up step of a short bench run (one full step = 3260 launches), -s skips the peak measurements
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:emagls\|oz -c 3400 \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-render --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_under_ncu.json 2>&1
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.summary.txt 2>&1
head -14 gpurun_out/${TAG}_launches.summary.txt
rm -f gpurun_out/${TAG}_launches.csv.gz; gzip -f gpurun_out/${TAG}_launches.csv
ls -la gpurun_out/ | grep ${TAG}
