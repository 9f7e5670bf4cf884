#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r01_v13}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --no-render --no-cpu-baseline > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_short.json"))
r=d["roofline"]
print(d["value"], d["e2e"]["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"])
print({k: round(v*d["ms_per_step"],1) for k,v in r["class_time_share"].items()})
PY
tail -5 gpurun_out/${TAG}_bench.err
cap() { # name regex skip count extra-env
  env $5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/$1 \
      python tools/gpu_ncu_factor.py 3600 > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
cap ${TAG}_sweep "gram_sweep_kernel" 1 1 A=1
cap ${TAG}_chol "gram_chol_kernel" 1 1 EMAGLS_GRAM_CHOL=1
grep -n "Duration\|Registers Per\|Achieved Occ\|Issue Slots Busy\|Executed Ipc Active\|L1/TEX Hit\|Mem Busy\|Bank\|bank" gpurun_out/${TAG}_sweep.details.txt gpurun_out/${TAG}_chol.details.txt | head -40
