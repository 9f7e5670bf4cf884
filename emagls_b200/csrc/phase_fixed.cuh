// Phase continuation t = |H_k| y / |y| (lib/getEMagLs2Filters.m:95-103) and the balanced base-256 digit split of t
// WITHOUT FP64 instructions, for the epilogue of the forward int8 tensor-core product (ozaki_kernels.cu).
//
// Why: on this part FP64 instructions do not overlap with tcgen05 MMAs (profiles/r02_v37_fwd_interference.txt: an FP64
// functor adds its full time to the launch, an integer functor of the same length vanishes behind the MMAs of the next
// tile).  Everything below runs on the integer pipes; the only floating-point instructions are one FP32 FMA / MUL
// pair and MUFU.RSQ for a 22-bit seed of 1 / |y|.
//
//   inputs   vr, vi     the exact integer results of the re / im column (phase_fixed_i64: the 64-bit integers of the
//                       integer drain; phase_fixed: FP64 bit patterns of the FP64 drain, rounded to 53 bits),
//            dexp       exponent(scale of the re row of c) - exponent(scale of the im row) (both powers of two),
//            mu_fix     |H_k| 2^(6 - e_H) 256^(T-4) 2^40 as an integer (< 2^62)
//   outputs  Zr, Zi  =  rn(t 2^24) with t = mu y / |y|,  |Z| <= 2^46: the integer whose balanced base-256 digits
//            are the int8 planes the backward product consumes.
//
//   A, B      |re|, |im| aligned to a common exponent, the larger one in [2^61, 2^62)
//   R0        2^54 rsqrt(a^2 + b^2) from the top 24 bits a, b of A, B (relative error eps0 <= 2^-21.5: 2 ulp of MUFU.RSQ,
//             the truncation of A, B and one FP32 rounding; the correction below holds up to 2^-20)
//   P, Q      A R0 / 2^32, B R0 / 2^32 = 2^60 (A, B) / |y| (1 + eps0)
//   e         1 - (P^2 + Q^2) / 2^120 = -2 eps0 + ...  (exact to 2^-54 from 32 x 32 -> 64 bit partial products)
//   U, W      P, Q times (1 + e/2 + 3 e^2 / 8): third order, remainder 5 e^3 / 16 < 2^-62
//   Z         (mu_fix U) / 2^76 rounded
// Error of Z against the exact value: <= 0.5 + 0.1 units (tests/test_phase_fixed.py runs this header on the CPU
// against a quad-precision evaluation, with the seed perturbed by the worst case of MUFU.RSQ).
#pragma once
#include <stdint.h>
#if !defined(__CUDA_ARCH__)
#include <math.h>
#include <string.h>
#endif

#if defined(__CUDACC__)
#define EM_PFX_HD __host__ __device__ __forceinline__
#else
#define EM_PFX_HD inline
#endif

namespace emagls {
namespace pfx {

#ifndef EM_PFX_SEED_PERTURB
#define EM_PFX_SEED_PERTURB 0.0f   // host tests: relative perturbation of the seed
#endif

EM_PFX_HD void dbits(double x, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  hi = (uint32_t)__double2hiint(x);
  lo = (uint32_t)__double2loint(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  hi = (uint32_t)(b >> 32);
  lo = (uint32_t)b;
#endif
}

// 53-bit significand (leading bit 52) and the biased exponent field of a normal double.  For zero (e == 0) the
// significand is NOT cleared: callers shift it out through the exponent (a shift by 63 of a value below 2^62).
EM_PFX_HD uint64_t significand(uint32_t hi, uint32_t lo, int& e) {
  e = (int)((hi >> 20) & 0x7FFu);
  return ((uint64_t)((hi & 0xFFFFFu) | 0x100000u) << 32) | lo;
}

EM_PFX_HD uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * (uint64_t)b; }   // IMAD.WIDE.U32

// x 2^(40 + 8 (T - 4)) up as an integer: x = |H_k| (>= 0), up = 2^(6 - e_H) with x up < 64
template <int T>
EM_PFX_HD uint64_t magnitude_fixed(double x, double up) {
  uint32_t xh, xl, uh, ul;
  dbits(x, xh, xl);
  dbits(up, uh, ul);
  int ex;
  const uint64_t m = significand(xh & 0x7FFFFFFFu, xl, ex);
  const int eu = (int)((uh >> 20) & 0x7FFu);
  // x up 2^(40 + 8 (T-4)) = m 2^(ex - 1075) 2^(eu - 1023) 2^(40 + 8 (T-4))
  const int sh = ex + eu - 2058 + 8 * (T - 4);
  if (sh >= 0) return m << (sh > 10 ? 10 : sh);      // x up < 64 implies sh <= 1 + 8 (T-4) - ... <= 10
  const int r = -sh;
  return r > 63 ? 0ull : (m >> r);
}

EM_PFX_HD uint32_t seed_r0(uint32_t a, uint32_t b) {
  const float af = (float)a, bf = (float)b;          // exact: a, b < 2^24
  const float nf = fmaf(af, af, bf * bf);            // in [2^46, 2^49)
#if defined(__CUDA_ARCH__)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(nf));   // MUFU.RSQ without the subnormal guards
  return __float2uint_rz(r * 18014398509481984.0f);           // 2^54 / sqrt(n): (2^29.5, 2^31]
#else
  float rf = (1.0f / sqrtf(nf)) * 18014398509481984.0f;
  rf *= (1.0f + EM_PFX_SEED_PERTURB);
  return (uint32_t)rf;
#endif
}

// one Newton-like third-order step shared by both components: returns C = (e/2 + 3 e^2 / 8) 2^57
EM_PFX_HD int64_t correction(uint64_t P, uint64_t Q) {
  const uint32_t Ph = (uint32_t)(P >> 32), Pl = (uint32_t)P, Qh = (uint32_t)(Q >> 32), Ql = (uint32_t)Q;
  // floor(P^2 / 2^64) up to 2 units: Ph^2 + (2 Ph Pl) / 2^32
  const uint64_t SP = mul_wide(Ph, Ph) + (mul_wide(Ph, Pl) >> 31);
  const uint64_t SQ = mul_wide(Qh, Qh) + (mul_wide(Qh, Ql) >> 31);
  const int64_t E56 = (int64_t)((1ull << 56) - SP - SQ);            // e 2^56, |E56| < 2^36
  const uint32_t eh = (uint32_t)((uint64_t)E56 >> 22);               // low word of E56 / 2^22, |.| < 2^15 (eps0 < 2^-20)
  const uint32_t sq = eh * eh;                                       // the square of a two's complement word
  return E56 + (int64_t)((3u * sq) >> 14);                           // 3 e^2 / 8 in units of 2^-57
}

EM_PFX_HD uint64_t apply_correction(uint64_t P, int64_t C) {
  const int32_t Ps = (int32_t)(P >> 31);                             // < 2^30 + 2^9
  const int32_t Cs = (int32_t)(C >> 6);                              // |C| < 2^37 for eps0 < 2^-20: |Cs| < 2^31
  const int64_t d = ((int64_t)Ps * (int64_t)Cs) >> 20;               // P C / 2^57
  return P + (uint64_t)d;
}

EM_PFX_HD uint64_t mul_hi_approx(uint32_t Mh, uint32_t Ml, uint64_t U) {   // floor(M U / 2^64) up to 3 units
  const uint32_t Uh = (uint32_t)(U >> 32), Ul = (uint32_t)U;
  return mul_wide(Mh, Uh) + (mul_wide(Mh, Ul) >> 32) + (mul_wide(Ml, Uh) >> 32);
}

// the normalisation on aligned magnitudes A, B (the larger one in [2^61, 2^62)) and the signs of re / im
EM_PFX_HD void phase_from_aligned(uint64_t A, uint64_t B, bool neg_r, bool neg_i, uint64_t mu_fix, int64_t& Zr, int64_t& Zi);

EM_PFX_HD void phase_fixed(double vr, double vi, int dexp, uint64_t mu_fix, int64_t& Zr, int64_t& Zi) {
  uint32_t rh, rl, ih, il;
  dbits(vr, rh, rl);
  dbits(vi, ih, il);
  int er, ei;
  const uint64_t Mr = significand(rh, rl, er), Mi = significand(ih, il, ei);
  if ((er | ei) == 0) {                                              // angle(0) = 0
    Zr = (int64_t)((mu_fix + 0x8000ull) >> 16);
    Zi = 0;
    return;
  }
  const int ear = er ? er + dexp : -(1 << 20), eai = ei ? ei : -(1 << 20);
  const int E = ear > eai ? ear : eai;
  const int sa = E - ear, sb = E - eai;                              // one of them is 0
  const uint64_t A = (Mr << 9) >> (sa > 63 ? 63 : sa), B = (Mi << 9) >> (sb > 63 ? 63 : sb);   // < 2^62
  phase_from_aligned(A, B, (rh >> 31) != 0, (ih >> 31) != 0, mu_fix, Zr, Zi);
}

// The same from the exact 64-bit integers (|v| < 2^63) the accumulator drain can deliver without FP64 instructions:
// no rounding to 53 bits in between.
EM_PFX_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
EM_PFX_HD void phase_fixed_i64(int64_t vr, int64_t vi, int dexp, uint64_t mu_fix, int64_t& Zr, int64_t& Zi) {
  if ((vr | vi) == 0) {                                              // angle(0) = 0
    Zr = (int64_t)((mu_fix + 0x8000ull) >> 16);
    Zi = 0;
    return;
  }
  const uint64_t ur = (uint64_t)(vr < 0 ? -vr : vr), ui = (uint64_t)(vi < 0 ? -vi : vi);
  const int nr = vr ? clz64(ur) : 0, ni = vi ? clz64(ui) : 0;        // >= 1 for nonzero values
  const int ear = vr ? dexp - nr : -(1 << 20), eai = vi ? -ni : -(1 << 20);   // exponent of the leading bit, minus 63
  const int E = ear > eai ? ear : eai;
  const int sa = E - ear, sb = E - eai;                              // one of them is 0
  const uint64_t A = ((ur << nr) >> 2) >> (sa > 63 ? 63 : sa), B = ((ui << ni) >> 2) >> (sb > 63 ? 63 : sb);   // < 2^62
  phase_from_aligned(A, B, vr < 0, vi < 0, mu_fix, Zr, Zi);
}

EM_PFX_HD void phase_from_aligned(uint64_t A, uint64_t B, bool neg_r, bool neg_i, uint64_t mu_fix, int64_t& Zr, int64_t& Zi) {
  const uint32_t R0 = seed_r0((uint32_t)(A >> 38), (uint32_t)(B >> 38));
  const uint64_t P = mul_wide((uint32_t)(A >> 32), R0) + (mul_wide((uint32_t)A, R0) >> 32);
  const uint64_t Q = mul_wide((uint32_t)(B >> 32), R0) + (mul_wide((uint32_t)B, R0) >> 32);
  const int64_t C = correction(P, Q);
  const uint64_t U = apply_correction(P, C), W = apply_correction(Q, C);
  const uint32_t Mh = (uint32_t)(mu_fix >> 32), Ml = (uint32_t)mu_fix;
  const int64_t zr = (int64_t)((mul_hi_approx(Mh, Ml, U) + 2048ull) >> 12);
  const int64_t zi = (int64_t)((mul_hi_approx(Mh, Ml, W) + 2048ull) >> 12);
  Zr = neg_r ? -zr : zr;
  Zi = neg_i ? -zi : zi;
}

// Balanced base-256 digits of Z = a 2^24 256^(T-4), |a| <= 64, as two words: byte j of zl is digit T-1-j (j < 3),
// byte j of zh is digit T-4-j (j < T-3)  (the layout of oz::slice_words)
template <int T>
EM_PFX_HD void split_words(int64_t Z, uint32_t& zl, uint32_t& zh) {
  static_assert(T >= 4 && T <= 6, "hi limb: 1..3 digits, lo limb: 3 digits");
  constexpr int nh = T - 3;
  constexpr int bias_hi = (nh == 3) ? 0x808080 : (nh == 2 ? 0x8080 : 0x80);
  int32_t hi = (int32_t)((Z + (1 << 23)) >> 24);                     // |hi| <= 2^22
  const int32_t lo = (int32_t)((uint32_t)Z - ((uint32_t)hi << 24));  // in [-2^23, 2^23)
  const int32_t yl = lo + 0x808080;                                  // in (0, 2^25)
  hi += yl >> 24;                                                    // carry (0 or 1)
  zl = (uint32_t)(yl ^ 0x808080);
  zh = (uint32_t)((hi + bias_hi) ^ bias_hi);
}

}  // namespace pfx
}  // namespace emagls
