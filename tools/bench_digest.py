"""Print the parts of a bench.py JSON line that matter when reading a gpurun tail.
usage: python tools/bench_digest.py gpurun_out/xxx_bench.json"""
import json
import sys

j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = j["roofline"]
e = j.get("e2e") or {}
print("BENCH", round(j["value"], 1), j["unit"], round(j["ms_per_step"], 1), "ms  e2e", round(e.get("value", 0), 1),
      "host_api", (e.get("host_api_call") or {}).get("value"), "diff", e.get("max_rel_diff_vs_device_resident_banks"),
      "launches", j.get("gpu_launches"), "clocks", j.get("clocks"))
print("dominant", r.get("dominant_class"), "frac", round(r["frac"], 3), "sweeps", r.get("jacobi_mean_sweeps"))
for k, c in (r.get("classes") or {}).items():
    print(f"  {k:10s} {c['ms_per_step']:8.1f} ms share {c['share']}  {c['achieved']:.2f}/{c['peak']:.1f} {c['unit']} frac {c['frac']:.3f}")
print("shares", r.get("class_time_share"))
for key in ("parity_spot_check", "cpu_baseline", "config3", "strong_scaling"):
    print(key, j.get(key))
rd = j.get("render") or {}
print("render", rd.get("value"), (rd.get("roofline") or {}).get("frac"), rd.get("e2e"))
