"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/xxx.txt"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, us))
agg = collections.OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot/1e3:.2f} ms total (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'launches':>9s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:9d} {us/1e3:10.3f} {us/c:10.2f} {us/tot:7.3f}")
