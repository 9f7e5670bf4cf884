#!/bin/bash
# ncu --set full captures of the hot kernels (run under gpurun).  `-k regex:` matches the function
# name without template arguments: ozaki_gemm_kernel launches alternate forward (EpiPhaseSlice) /
# backward (EpiStoreF64).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01_v8}
cap() { # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
      python tools/gpu_ncu_factor.py 3600 > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 15000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
cap ${TAG}_oz "ozaki_gemm_kernel" 40 2
cap ${TAG}_factor "factor_kernel" 2 1
ls -la gpurun_out/
