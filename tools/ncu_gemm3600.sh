#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 9 -c 2 -f -o gpurun_out/r01_v6_gemm3600 \
    python tools/gpu_ncu_factor.py 3600 > gpurun_out/r01_v6_gemm3600.log 2>&1
ncu -i gpurun_out/r01_v6_gemm3600.ncu-rep --page details > gpurun_out/r01_v6_gemm3600.details.txt 2>/dev/null
ncu -i gpurun_out/r01_v6_gemm3600.ncu-rep --page raw --csv > gpurun_out/r01_v6_gemm3600.raw.csv 2>/dev/null
rm -f gpurun_out/r01_v6_gemm3600.ncu-rep
tail -2 gpurun_out/r01_v6_gemm3600.log
