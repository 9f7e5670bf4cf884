"""SASS evidence for the hand-written kernels: per kernel family, the instructions that prove the hardware path
(tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, FP64 tensor core -> DMMA, cp.async -> LDGSTS,
mbarrier -> SYNCS).  usage: python tools/sass_summary.py > profiles/r01_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "emagls_b200", "lib", "libemagls_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
WANT = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|LDTM|UTMALDG|UTMASTG|UBLKCP|DMMA|HMMA|LDGSTS|SYNCS|DFMA|DMUL|DADD|LDG|STG|LDS|STS|SHFL|BAR)\b")
fams = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        base = re.sub(r"^void ", "", name).split("(")[0]
        if "gemm_f64_kernel" in base:
            base = "emagls::gemm_f64_kernel<tile config, epilogue, operand layouts> (all instances)"
        base = re.sub(r"<([^<>]|<[^<>]*>)*>$", lambda mm: mm.group(0) if len(mm.group(0)) < 60 else "<..>", base)
        cur = fams.setdefault(base, {"instances": 0, "instr": 0, "ops": collections.Counter()})
        cur["instances"] += 1
        continue
    if cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        cur["instr"] += 1
        m = WANT.search(line)
        if m:
            cur["ops"][m.group(1)] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): static instruction counts, summed over the template instances")
print(f"{'kernel family':78s} {'inst':>4s} {'instr':>7s}  evidence")
for base, v in sorted(fams.items(), key=lambda kv: -kv[1]["instr"]):
    key = [k for k in ("UTCIMMA", "UTCQMMA", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "DMMA", "HMMA", "LDGSTS", "SYNCS") if v["ops"].get(k)]
    key += [k for k in v["ops"] if k.startswith("UTC") and k not in key]
    rest = [k for k in ("DFMA", "DMUL", "DADD", "SHFL", "LDS", "STS", "LDG", "STG", "BAR") if v["ops"].get(k)]
    print(f"{base[:78]:78s} {v['instances']:4d} {v['instr']:7d}  " + " ".join(f"{k}={v['ops'][k]}" for k in key + rest))
