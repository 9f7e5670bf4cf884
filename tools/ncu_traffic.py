"""Extract per-launch DRAM traffic / duration / tensor-pipe activity of captured kernels from an
`ncu --page raw --csv` export and write the small JSON bench.py reads for roofline.traffic.
usage: python tools/ncu_traffic.py gpurun_out/x.raw.csv profiles/r01_gemm_traffic.json "<note>" """
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
def col(name):
    i = hdr.index(name)
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "%": 1.0}.get(units[i], 1.0)
    return [float(r[i].replace(",", "")) * scale for r in rows[2:]]
out = []
names = [r[hdr.index("Kernel Name")] for r in rows[2:]]
rd, wr, dur = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
tp = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
grid = [r[hdr.index("launch__grid_size")] for r in rows[2:]]
for n, a, b, d, t, g in zip(names, rd, wr, dur, tp, grid):
    out.append({"kernel": n.split("(")[0], "grid": int(g), "dram_read_bytes": a, "dram_write_bytes": b,
                "traffic_bytes": a + b, "duration_s_under_ncu": d, "tensor_pipe_active_pct": t})
json.dump({"note": sys.argv[3] if len(sys.argv) > 3 else "", "launches": out}, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
