#!/bin/bash
# One iteration on a B200: GPU parity tests (all, or the design/variant/block/arbitration files with "quick"), a short
# bench line per extra "ENV=value" argument (A=1 = defaults), optionally the launch list of one 3600-orientation design
# ("launches") and --set full captures ("cap:<name>:<regex>:<skip>").  usage: tools/gpu_r2_step.sh <tag> [quick] [launches] [cap:...] [ENV=v ...]
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
TESTS="tests"
for a in "$@"; do
  case "$a" in
    quick) TESTS="tests/test_gpu_design.py tests/test_gpu_variants.py tests/test_gpu_blocks.py tests/test_gpu_arbitration.py";;
  esac
done
timeout 1500 python -m pytest $TESTS -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
for a in "$@"; do
  case "$a" in
    quick) ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
          python tools/gpu_ncu_factor.py 3600 > gpurun_out/${TAG}_launches.log 2>&1
      python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.summary.txt 2>&1
      head -30 gpurun_out/${TAG}_launches.summary.txt
      rm -f gpurun_out/${TAG}_launches.csv.gz; gzip -f gpurun_out/${TAG}_launches.csv;;
    cap:*)
      IFS=: read _ name regex skip <<< "$a"
      out=gpurun_out/${TAG}_${name}
      timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s ${skip:-1} -c 1 -f -o $out \
          python tools/gpu_ncu_factor.py 3600 > $out.log 2>&1
      ncu -i $out.ncu-rep --page details > $out.details.txt 2>/dev/null
      ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
      ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
      echo "== $name"; grep -E "^  [a-z].*\(|Duration|Registers Per|Achieved Occ|Executed Ipc Active|No Eligible|Mem Busy|Max Bandwidth|DRAM Throughput|L2 Hit|L1/TEX Hit|Bank|cycles being stalled" $out.details.txt | head -30
      sz=$(stat -c %s $out.ncu-rep 2>/dev/null || echo 0)
      if [ "$sz" -gt 9000000 ]; then rm -f $out.ncu-rep; fi;;
    *=*)
      env $a timeout 600 python bench.py --steps 3 --warmup 3 --no-render --no-cpu-baseline --no-spot-check > gpurun_out/${TAG}_bench_${a//[^A-Za-z0-9_]/_}.json 2>> gpurun_out/${TAG}_bench.err
      python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_${a//[^A-Za-z0-9_]/_}.json").read().strip().splitlines()[-1])
    print("AB ${a}", round(j["value"], 1), round(j["ms_per_step"], 1), {k: round(v["ms_per_step"], 1) for k, v in j["roofline"]["classes"].items()})
except Exception as e:
    print("ab parse failed", repr(e)); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
      ;;
  esac
done
