#!/bin/bash
# SASS evidence for the hand-written kernels: per kernel, the count of the instructions that prove the
# hardware path (tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld -> LDTM, DMMA, cp.async -> LDGSTS,
# mbarrier -> SYNCS) plus registers / spills from ptxas.  usage: tools/sass_summary.sh > profiles/r01_sass_summary.txt
cd "$(dirname "$0")/.."
LIB=emagls_b200/lib/libemagls_cuda.so
echo "# cuobjdump -sass $LIB (sm_100a), instruction counts per kernel"
cuobjdump -sass $LIB 2>/dev/null | awk '
  /Function : /{ if (fn != "") print_fn(); fn=$3; delete c; n=0 }
  /^[ \t]+\/\*[0-9a-f]+\*\// { n++; for (i=1;i<=NF;i++) { if ($i ~ /^(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|LDTM|UTCBAR|DMMA|LDGSTS|SYNCS|DFMA|STG|LDG)(\.|$|;)/) { split($i, a, "."); sub(/;/, "", a[1]); c[a[1]]++ } } }
  function print_fn(   k, s) { s=""; for (k in c) s = s " " k "=" c[k]; printf "%-110s instr=%d%s\n", substr(fn,1,110), n, s }
  END { if (fn != "") print_fn() }' | c++filt 2>/dev/null | grep -E "UTC|UTMALDG|DMMA|LDGSTS|factor_kernel|gram_sweep|bwd_small|fused_render|spectral_mac|channel_mix" | sort
