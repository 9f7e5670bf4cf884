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

// apply Q_C (forward = false: Q_C * x, blocks/reflectors in reverse order with tau) or
// Q_C^H (forward = true: blocks/reflectors in order with conj(tau)) to x0, x1 (length S)
__device__ inline void apply_qc(const BlockPlan& bp, const cplx* __restrict__ V, const cplx* __restrict__ tau,
                         cplx* x0, cplx* x1, bool adjoint, int lane) {
  const int S = bp.S, Mc = bp.Mc;
  for (int tt = 0; tt < bp.nblk; ++tt) {
    const int t = adjoint ? tt : bp.nblk - 1 - tt;
    const int r0 = (t == 0) ? 0 : bp.R0 + (t - 1) * bp.RB;
    const int r1 = (t == 0) ? bp.R0 : min(S, r0 + bp.RB);
    for (int jj = 0; jj < Mc; ++jj) {
      const int j = adjoint ? jj : Mc - 1 - jj;
      cplx ta = tau[t * bp.MC + j];
      if (ta.x == 0.0 && ta.y == 0.0) continue;
      if (adjoint) ta.y = -ta.y;
      const int lo = (t == 0) ? j + 1 : r0;
      const cplx* v = V + (long long)j * S;
      cplx w0 = mk(0.0, 0.0), w1 = mk(0.0, 0.0);
      if (lane == 0) { w0 = x0[j]; w1 = x1[j]; }
      for (int i = lo + lane; i < r1; i += 32) {
        cplx vi = v[i];
        cfmac(w0, vi, x0[i]);
        cfmac(w1, vi, x1[i]);
      }
      w0 = wsumc(w0); w1 = wsumc(w1);
      cplx f0 = cmul(ta, w0), f1 = cmul(ta, w1);
      if (lane == 0) { x0[j] = csub(x0[j], f0); x1[j] = csub(x1[j], f1); }
      for (int i = lo + lane; i < r1; i += 32) {
        cplx vi = v[i];
        cplx a0 = x0[i], a1 = x1[i];
        cfms(a0, f0, vi); cfms(a1, f1, vi);
        x0[i] = a0; x1[i] = a1;
      }
      __syncwarp();
    }
  }
}

}  // namespace emagls
