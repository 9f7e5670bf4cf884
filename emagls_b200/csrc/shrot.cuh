// Real SH rotation matrices by the Ivanic-Ruedenberg band recursion
// (dependencies/Spherical-Harmonic-Transform/getSHrotMtx.m:59-121), shared by the EMA-in-SH steering
// rows (ema_kernels.cu) and the SH-signal rotation of the render front end (frontend.cu).
#pragma once
#include "common.cuh"

namespace emagls {

// R_1 band of getSHrotMtx (index = m + 1 for m in (-1,0,1) <-> (y,z,x)) from the 3x3 matrix E (row-major)
struct Rot1 { double r[3][3]; };

__device__ __forceinline__ double rot_P(const Rot1& R1, const double* Rlm1, int ldl, int i, int l, int a, int b) {
  // Rlm1: band l-1, [(2l-1) x (2l-1)], row a + l - 1
  const double ri1 = R1.r[i + 1][2], rim1 = R1.r[i + 1][0], ri0 = R1.r[i + 1][1];
  const double* row = Rlm1 + (a + l - 1) * ldl;
  if (b == -l) return ri1 * row[0] + rim1 * row[2 * l - 2];
  if (b == l) return ri1 * row[2 * l - 2] - rim1 * row[0];
  return ri0 * row[b + l - 1];
}

// R_1 from the 3x3 rotation matrix E (row-major): rows/columns ordered (y, z, x)  (getSHrotMtx.m:70-74)
__device__ __forceinline__ void rot1_from_matrix(const double E[3][3], Rot1& R1) {
  R1.r[0][0] = E[1][1]; R1.r[0][1] = E[1][2]; R1.r[0][2] = E[1][0];
  R1.r[1][0] = E[2][1]; R1.r[1][1] = E[2][2]; R1.r[1][2] = E[2][0];
  R1.r[2][0] = E[0][1]; R1.r[2][1] = E[0][2]; R1.r[2][2] = E[0][0];
}

// Cooperative over the nt threads of one CTA (contains __syncthreads): Rr [nsh][nsh] row-major,
// nsh = (order+1)^2, block diagonal.  R1 and *rotate_flag live in shared memory and may have been
// written by one thread just before the call: they are read only after the first barrier.
// *rotate_flag == 0 leaves the identity.
__device__ inline void sh_rot_real_bands(const Rot1& R1, const int* rotate_flag, int order, double* Rr, int tid,
                                         int nt) {
  const int nsh = (order + 1) * (order + 1);
  for (int i = tid; i < nsh * nsh; i += nt) Rr[i] = (i / nsh == i % nsh) ? 1.0 : 0.0;
  __syncthreads();
  if (*rotate_flag && order >= 1) {
    for (int i = tid; i < nsh * nsh; i += nt) Rr[i] = (i == 0) ? 1.0 : 0.0;
    __syncthreads();
    if (tid < 9) Rr[(1 + tid / 3) * nsh + 1 + tid % 3] = R1.r[tid / 3][tid % 3];
    __syncthreads();
    for (int l = 2; l <= order; ++l) {
      const int w = 2 * l + 1, band = l * l, bandm1 = (l - 1) * (l - 1);
      const double* Rl1 = Rr + bandm1 * nsh + bandm1;   // band l-1 block, leading dimension nsh
      for (int e = tid; e < w * w; e += nt) {
        const int m = e / w - l, n = e % w - l;
        const int am = m < 0 ? -m : m, an = n < 0 ? -n : n;
        const double dm = (m == 0) ? 1.0 : 0.0;
        const double denom = (an == l) ? (double)(2 * l) * (2 * l - 1) : (double)(l * l - n * n);
        double u = sqrt((double)(l * l - m * m) / denom);
        double v = sqrt((1.0 + dm) * (double)(l + am - 1) * (double)(l + am) / denom) * (1.0 - 2.0 * dm) * 0.5;
        double ww = sqrt((double)(l - am - 1) * (double)(l - am) / denom) * (1.0 - dm) * (-0.5);
        if (u != 0.0) u *= rot_P(R1, Rl1, nsh, 0, l, m, n);
        if (v != 0.0) {
          double V;
          if (m == 0) V = rot_P(R1, Rl1, nsh, 1, l, 1, n) + rot_P(R1, Rl1, nsh, -1, l, -1, n);
          else if (m > 0) {
            const double dd = (m == 1) ? 1.0 : 0.0;
            V = rot_P(R1, Rl1, nsh, 1, l, m - 1, n) * sqrt(1.0 + dd);
            if (dd == 0.0) V -= rot_P(R1, Rl1, nsh, -1, l, -m + 1, n);
          } else {
            const double dd = (m == -1) ? 1.0 : 0.0;
            V = rot_P(R1, Rl1, nsh, -1, l, -m - 1, n) * sqrt(1.0 + dd);
            if (dd == 0.0) V += rot_P(R1, Rl1, nsh, 1, l, m + 1, n);
          }
          v *= V;
        }
        if (ww != 0.0) {
          double Wv;
          if (m > 0) Wv = rot_P(R1, Rl1, nsh, 1, l, m + 1, n) + rot_P(R1, Rl1, nsh, -1, l, -m - 1, n);
          else Wv = rot_P(R1, Rl1, nsh, 1, l, m - 1, n) - rot_P(R1, Rl1, nsh, -1, l, -m + 1, n);
          ww *= Wv;
        }
        Rr[(band + m + l) * nsh + band + n + l] = u + v + ww;
      }
      __syncthreads();
    }
  }
}

}  // namespace emagls
