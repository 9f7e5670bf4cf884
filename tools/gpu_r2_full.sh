#!/bin/bash
# Full validation of a tree on one B200: all GPU parity tests, smoke(), the default bench line and the
# reference arm.  usage: tools/gpu_r2_full.sh <tag> [extra env assignments for an A/B short bench ...]
cd "$(dirname "$0")/.."
TAG=${1:-r02}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "^c1:|^c5:" gpurun_out/${TAG}_pytest_gpu.log | head -40
tail -6 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("BENCH", round(j["value"], 1), "sets/s", round(j["ms_per_step"], 1), "ms  e2e", round(j["e2e"]["value"], 1),
          "host_api", j["e2e"].get("host_api_call"), "diff", j["e2e"].get("max_rel_diff_vs_device_resident_banks"))
    print("dominant", r["dominant_class"], "frac", r["frac"], "sweeps", r.get("jacobi_mean_sweeps"))
    for k, c in r["classes"].items():
        print(f"  {k:10s} {c['ms_per_step']:8.1f} ms share {c['share']}  {c['achieved']:.2f}/{c['peak']:.1f} {c['unit']} frac {c['frac']:.3f}")
    print("spot", j.get("parity_spot_check"))
    print("cpu", j.get("cpu_baseline"))
    print("render", (j.get("render") or {}).get("value"), ((j.get("render") or {}).get("roofline") or {}).get("frac"))
    print("config3", j.get("config3"))
    print("strong", j.get("strong_scaling"))
except Exception as e:
    print("bench parse failed", repr(e)); print(open("gpurun_out/${TAG}_bench.err").read()[-3000:])
PY
for kv in "$@"; do
  env $kv timeout 600 python bench.py --steps 3 --warmup 3 --no-render --no-cpu-baseline --no-spot-check > gpurun_out/${TAG}_bench_${kv//[^A-Za-z0-9_]/_}.json 2>> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_${kv//[^A-Za-z0-9_]/_}.json").read().strip().splitlines()[-1])
    print("AB ${kv}", round(j["value"], 1), round(j["ms_per_step"], 1), j["roofline"]["class_time_share"])
except Exception as e:
    print("ab parse failed", repr(e))
PY
done
