// CPU check of emagls_b200/csrc/phase_fixed.cuh (the FP64-free phase continuation + digit split of the forward
// tensor-core product): the header is compiled for the host and compared with a quad-precision evaluation of
// t = |H| y / |y|.  Prints the largest error of Z = rn(t 2^24) in units of Z and the number of digit failures.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include <quadmath.h>
#include "../../emagls_b200/csrc/phase_fixed.cuh"

using namespace emagls::pfx;

template <int T>
static int run(long long n, unsigned seed, double& worst) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  int bad = 0;
  for (long long it = 0; it < n; ++it) {
    // integer-valued doubles of widely varying size, zeros, equal magnitudes, one dominant component
    auto rnd_int = [&](int kind) -> double {
      if (kind == 0) return 0.0;
      const int bits = 1 + (int)(uni(rng) * 70);
      double v = std::floor(std::ldexp(uni(rng) + 1.0, bits - 1));
      if (bits > 53) v = std::ldexp(std::floor(std::ldexp(uni(rng) + 1.0, 52)), bits - 53);
      return (rng() & 1) ? -v : v;
    };
    const int kr = (it % 97 == 0) ? 0 : 1, ki = (it % 89 == 0) ? 0 : 1;
    double vr = rnd_int(kr), vi = rnd_int(ki);
    if (it % 13 == 0 && vr != 0.0) vi = (rng() & 1) ? vr : -vr;
    if (it % 17 == 0) vi = (rng() & 1) ? 1.0 : -1.0;
    const int dexp = (it % 5 == 0) ? 0 : (int)(uni(rng) * 81) - 40;
    const int eH = (int)(uni(rng) * 40) - 20;                       // 2^eH > max |H|
    double x = std::ldexp(uni(rng), eH);                            // |H| in [0, 2^eH)
    if (it % 7 == 0) x = std::ldexp(1.0 - std::ldexp(1.0, -53), eH);   // the largest admissible magnitude
    if (it % 101 == 0) x = std::ldexp(uni(rng), eH - 45);           // tiny against the row maximum
    const double up = std::ldexp(1.0, 6 - eH);
    const uint64_t mu_fix = magnitude_fixed<T>(x, up);
    int64_t Zr, Zi;
    phase_fixed(vr, vi, dexp, mu_fix, Zr, Zi);
    const __float128 scale = (__float128)x * (__float128)up * ldexpq((__float128)1.0, 24 + 8 * (T - 4));
    __float128 er, ei;
    if (vr == 0.0 && vi == 0.0) { er = scale; ei = 0; }
    else {
      const __float128 re = ldexpq((__float128)vr, dexp), im = (__float128)vi;
      const __float128 nrm = sqrtq(re * re + im * im);
      er = scale * re / nrm; ei = scale * im / nrm;
    }
    const double dr = (double)fabsq((__float128)Zr - er), di = (double)fabsq((__float128)Zi - ei);
    worst = std::fmax(worst, std::fmax(dr, di));
    if (std::fabs(vr) < 9.2e18 && std::fabs(vi) < 9.2e18) {         // the entry on exact 64-bit integers
      int64_t Yr, Yi;
      phase_fixed_i64((int64_t)vr, (int64_t)vi, dexp, mu_fix, Yr, Yi);
      const double d2 = std::fmax((double)fabsq((__float128)Yr - er), (double)fabsq((__float128)Yi - ei));
      worst = std::fmax(worst, d2);
      if (std::llabs(Yr - Zr) > 1 || std::llabs(Yi - Zi) > 1) { if (bad < 5) printf("i64 / f64 entry disagree: %lld %lld vs %lld %lld\n", (long long)Yr, (long long)Yi, (long long)Zr, (long long)Zi); ++bad; }
    }
    for (int c = 0; c < 2; ++c) {
      const int64_t Z = c ? Zi : Zr;
      uint32_t zl, zh;
      split_words<T>(Z, zl, zh);
      int64_t sum = 0;
      bool ok = true;
      for (int j = 0; j < T; ++j) {                                 // digit T-1-j has weight 256^j
        const int d = (int)(signed char)(j < 3 ? (zl >> (8 * j)) : (zh >> (8 * (j - 3))));
        sum += (int64_t)d * ((int64_t)1 << (8 * j));
        if (j == T - 1 && std::abs(d) > 65) ok = false;
      }
      if (sum != Z || !ok) { if (bad < 5) printf("digit failure T=%d Z=%lld sum=%lld\n", T, (long long)Z, (long long)sum); ++bad; }
    }
  }
  return bad;
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 2000000;
  double w6 = 0.0, w4 = 0.0;
  const int b6 = run<6>(n, 1234u, w6), b4 = run<4>(n, 99u, w4);
  printf("T=6 worst %.6f bad %d\nT=4 worst %.6f bad %d\n", w6, b6, w4, b4);
  return 0;
}
