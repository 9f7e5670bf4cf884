"""Per-source-line stall reasons from an `ncu --page source --csv` export (nvdisasm line info for the mapping).
usage: ncu_stalls.py <source.csv> <nvdisasm --print-line-info output> <kernel substring> [top n]"""
import csv, re, sys, collections
src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 20
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l and l.endswith(':'))
cur = None; amap = {}
for l in lines[start + 1:]:
    if l.startswith('.text.') and amap: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and cur: amap[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.defaultdict(lambda: collections.Counter()); base = None; tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[ci['Address']], 16) if r[ci['Address']].startswith('0x') else int(r[ci['Address']])
    if base is None: base = a
    key = amap.get(a - base, (('?', 0), ''))[0]
    for s in stalls:
        v = int(float(r[ci[s]] or 0)); agg[key][s] += v; tot += v
    agg[key]['_instr'] += int(float(r[ci['Instructions Executed']] or 0))
    agg[key]['_wf'] += int(float(r[ci['L1 Wavefronts Shared']] or 0))
    agg[key]['_wfi'] += int(float(r[ci['L1 Wavefronts Shared Ideal']] or 0))
allc = collections.Counter()
for k, c in agg.items():
    for s in stalls: allc[s] += c[s]
print("overall:", ", ".join(f"{s[6:]} {100*v/tot:.1f}%" for s, v in allc.most_common(8)))
for key, c in sorted(agg.items(), key=lambda kv: -sum(kv[1][s] for s in stalls))[:top]:
    t = sum(c[s] for s in stalls)
    best = ", ".join(f"{s[6:]} {100*c[s]/max(t,1):.0f}%" for s in sorted(stalls, key=lambda s: -c[s])[:3])
    print(f"{key[0]:>20s}:{key[1]:<5d} {100*t/tot:5.1f}%  instr {c['_instr']:>12d}  smem wf {c['_wf']}/{c['_wfi']}  | {best}")
