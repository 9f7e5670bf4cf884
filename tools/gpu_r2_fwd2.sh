#!/bin/bash
# Forward product, integer functor variants: microbenchmark (bytes of fixw* must equal fix), one --set full capture of
# the default variant inside the microbenchmark with the per-line stall samples.  usage: tools/gpu_r2_fwd2.sh <tag> [mode]
cd "$(dirname "$0")/.."
TAG=${1:-r02_fwd2}; MODE=${2:-8}
mkdir -p gpurun_out
{
  timeout 120 tools/microbench/bin/oz_fwd_bench 3600 1 20 6
  timeout 60 tools/microbench/bin/oz_fwd_bench 450 2 20 6
  timeout 60 tools/microbench/bin/oz_fwd_bench 450 2 20 4
  timeout 60 tools/microbench/bin/oz_fwd_bench 37 1 20 6
} 2>&1 | grep -v "mismatch" > gpurun_out/${TAG}_fwd_microbench.txt
grep -E "^oz_fwd|vs fix|dbg=" gpurun_out/${TAG}_fwd_microbench.txt
out=gpurun_out/${TAG}_fwd_cap
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:ozaki_gemm_kernel" -s 200 -c 1 -f -o $out \
    tools/microbench/bin/oz_fwd_bench 3600 1 20 6 $MODE > $out.log 2>&1
ncu -i $out.ncu-rep --page details > $out.details.txt 2>/dev/null
ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
rm -f $out.ncu-rep
grep -E "Duration|Registers Per|Executed Ipc Active|No Eligible|Issue Slots Busy|Local" $out.details.txt | head
