#!/bin/bash
# ncu --set full capture of selected kernels of one 3600-orientation design.  usage: tools/gpu_r2_ncu.sh <tag> <name:regex:skip> ...
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read name regex skip <<< "$spec"
  out=gpurun_out/${TAG}_${name}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$regex -s ${skip:-1} -c 1 -f -o $out \
      python tools/gpu_ncu_factor.py 3600 > $out.log 2>&1
  ncu -i $out.ncu-rep --page details > $out.details.txt 2>/dev/null
  ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
  ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
  grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|Executed Ipc Active|No Eligible|FP64|Shared Memory Configuration|Bank|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth|Local" $out.details.txt | head -40
  sz=$(stat -c %s $out.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 20000000 ]; then rm -f $out.ncu-rep; fi
done
