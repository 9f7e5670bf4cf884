"""Aggregate an `ncu --page source --csv` export by CUDA source line using nvdisasm line info.
usage: ncu_by_line.py <source.csv> <nvdisasm --print-line-info output> <mangled-kernel-substring> [kernel index in the csv]"""
import csv, re, sys, collections
src_csv, sass, kname = sys.argv[1:4]
# address -> (file, line) map from nvdisasm
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l and l.endswith(':'))
cur = None; amap = {}
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//--------------------- .text'):
        if amap: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        # inlined-at info may follow; keep innermost
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and cur:
        amap[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = his[which]
rows = rows[:his[which + 1]] if which + 1 < len(his) else rows
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0])
base = None
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[ci['Address']], 16) if r[ci['Address']].startswith('0x') else int(r[ci['Address']])
    if base is None: base = a
    off = a - base
    s = int(float(r[ci['# Samples']] or 0)); ins = int(float(r[ci['Instructions Executed']] or 0))
    key = amap.get(off, (('?', 0), ''))[0]
    agg[key][0] += s; agg[key][1] += ins
    tot_s += s; tot_i += ins
print(f"total samples {tot_s}, warp instructions {tot_i}")
for key, (s, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    print(f"{key[0]:>22s}:{key[1]:<5d} samples {100*s/tot_s:5.1f}%  instr {100*ins/max(tot_i,1):5.1f}%")
