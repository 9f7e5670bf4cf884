#!/bin/bash
# ncu captures of the v2 hot path (run under gpurun): per-kernel --set full reports + a launch list
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
cap r01_v2_gemm_fwd "EpiPhase" 0 1
cap r01_v2_gemm_bwd "gemm_f64_kernel.*EpiStore" 7 1
cap r01_v2_factor_lo "factor_kernel" 2 1
cap r01_v2_gram_chol "gram_chol" 0 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_launches_v2.csv \
    python tools/gpu_ncu_factor.py $B > gpurun_out/r01_launches_v2.log 2>&1
ls -la gpurun_out/
