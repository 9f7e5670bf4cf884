// Application of the stored Householder reflectors of the TSQR factorisation (solver_kernels.cu).
#pragma once
#include "kernels.h"

namespace emagls {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__device__ __forceinline__ cplx wsumc(cplx v) { v.x = wsum(v.x); v.y = wsum(v.y); return v; }
// Warp sums of two complex values at once: reduce-scatter over the first two butterfly stages (a lane keeps one of
// the four scalars, chosen by lane bits 4 and 3), three stages on that scalar, all-gather back: 18 shuffles instead of
// the 40 of two wsumc() calls (the applied-reflector chains are bound by the shuffle pipe: one warp instruction per
// clock and SM).  Every lane ends up with the same bits (the butterfly additions commute).
__device__ __forceinline__ void wsum2c(cplx& a, cplx& b, int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  cplx keep = hi16 ? b : a;
  const cplx send = hi16 ? a : b;
  keep.x += __shfl_xor_sync(0xffffffffu, send.x, 16);
  keep.y += __shfl_xor_sync(0xffffffffu, send.y, 16);
  double k = hi8 ? keep.y : keep.x;
  k += __shfl_xor_sync(0xffffffffu, hi8 ? keep.x : keep.y, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  const double o = __shfl_xor_sync(0xffffffffu, k, 8);       // the other component of the same value
  cplx c, d;
  c.x = hi8 ? o : k; c.y = hi8 ? k : o;                      // the complete sum of (hi16 ? b : a)
  d.x = __shfl_xor_sync(0xffffffffu, c.x, 16);
  d.y = __shfl_xor_sync(0xffffffffu, c.y, 16);
  a = hi16 ? d : c; b = hi16 ? c : d;
}

// apply Q_C (forward = false: Q_C * x, blocks/reflectors in reverse order with tau) or
// Q_C^H (forward = true: blocks/reflectors in order with conj(tau)) to x0, x1 (length S).
// The nblk * Mc reflectors are a strictly sequential chain (each needs the updated x), but the
// reflector vectors themselves do not depend on x: the next one is fetched from global memory into
// registers (<= RQ_NV values per lane) while the current one is applied, so the HBM/L2 latency of
// the 205 KB of reflectors per problem is hidden behind the two reductions of the step.
constexpr int RQ_NV = 6;   // ceil(max rows per block / 32): 128 rows (MC = 32), 192 rows (MC = 64)

struct ReflStep { int j, lo, r1; cplx ta; };

__device__ __forceinline__ ReflStep refl_step(const BlockPlan& bp, const cplx* __restrict__ tau, bool adjoint, int s) {
  const int Mc = bp.Mc;
  const int tt = s / Mc, jj = s - tt * Mc;
  const int t = adjoint ? tt : bp.nblk - 1 - tt;
  ReflStep st;
  st.j = adjoint ? jj : Mc - 1 - jj;
  const int r0 = (t == 0) ? 0 : bp.R0 + (t - 1) * bp.RB;
  st.r1 = (t == 0) ? bp.R0 : min(bp.S, r0 + bp.RB);
  st.lo = (t == 0) ? st.j + 1 : r0;
  st.ta = tau[t * bp.MC + st.j];
  return st;
}

__device__ inline void apply_qc(const BlockPlan& bp, const cplx* __restrict__ V, const cplx* __restrict__ tau,
                         cplx* x0, cplx* x1, bool adjoint, int lane) {
  const int S = bp.S, total = bp.nblk * bp.Mc;
  ReflStep cur = refl_step(bp, tau, adjoint, 0);
  cplx vc[RQ_NV], vn[RQ_NV];
#pragma unroll
  for (int u = 0; u < RQ_NV; ++u) {
    const int i = cur.lo + lane + 32 * u;
    vc[u] = (i < cur.r1) ? V[(long long)cur.j * S + i] : mk(0.0, 0.0);
  }
  for (int s = 0; s < total; ++s) {
    ReflStep nxt = cur;
    if (s + 1 < total) {
      nxt = refl_step(bp, tau, adjoint, s + 1);
#pragma unroll
      for (int u = 0; u < RQ_NV; ++u) {
        const int i = nxt.lo + lane + 32 * u;
        vn[u] = (i < nxt.r1) ? V[(long long)nxt.j * S + i] : mk(0.0, 0.0);
      }
    }
    cplx ta = cur.ta;
    if (ta.x != 0.0 || ta.y != 0.0) {
      if (adjoint) ta.y = -ta.y;
      const int j = cur.j;
      cplx w0 = mk(0.0, 0.0), w1 = mk(0.0, 0.0);
      if (lane == 0) { w0 = x0[j]; w1 = x1[j]; }
#pragma unroll
      for (int u = 0; u < RQ_NV; ++u) {
        const int i = cur.lo + lane + 32 * u;
        if (i < cur.r1) { cfmac(w0, vc[u], x0[i]); cfmac(w1, vc[u], x1[i]); }
      }
      wsum2c(w0, w1, lane);
      const cplx f0 = cmul(ta, w0), f1 = cmul(ta, w1);
      if (lane == 0) { x0[j] = csub(x0[j], f0); x1[j] = csub(x1[j], f1); }
#pragma unroll
      for (int u = 0; u < RQ_NV; ++u) {
        const int i = cur.lo + lane + 32 * u;
        if (i < cur.r1) {
          cplx a0 = x0[i], a1 = x1[i];
          cfms(a0, f0, vc[u]); cfms(a1, f1, vc[u]);
          x0[i] = a0; x1[i] = a1;
        }
      }
      __syncwarp();
    }
    cur = nxt;
#pragma unroll
    for (int u = 0; u < RQ_NV; ++u) vc[u] = vn[u];
  }
}

}  // namespace emagls
