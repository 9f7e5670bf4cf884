#!/bin/bash
# ncu captures of the v7 hot path (int8 tensor-core GEMMs); run under gpurun.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01_v7}
cap() { # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
      python tools/gpu_ncu_factor.py 3600 > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 15000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
cap ${TAG}_oz_fwd "ozaki_gemm_kernel.*EpiPhaseSlice" 20 1
cap ${TAG}_oz_bwd "ozaki_gemm_kernel.*EpiStoreF64" 20 1
cap ${TAG}_factor "factor_kernel" 2 1
cap ${TAG}_small "fwd_small_kernel|bwd_small_kernel|slice_rows_kernel|gram_chol_kernel|chain_bwd_kernel" 40 6
# launch list of the bench command itself (per-launch times are cold-cache and serialised: compare shares)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.summary.txt
gzip -f gpurun_out/${TAG}_launches.csv
ls -la gpurun_out/
