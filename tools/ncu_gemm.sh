#!/bin/bash
cd "$(dirname "$0")/.."
B=${1:-900}
cap() { # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
      python tools/gpu_ncu_factor.py $B > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 15000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
cap r01_v2_gemm "gemm_f64_kernel" 9 2
EMAGLS_DEBUG_INFO=1 python tools/gpu_ncu_factor.py 64 2> gpurun_out/r01_v2_debug_info.txt | tail -1
grep -E "^bin" gpurun_out/r01_v2_debug_info.txt | head -80
