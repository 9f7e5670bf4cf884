// Special functions on device: orthonormal spherical harmonics and spherical Bessel functions.
#pragma once
#include "common.cuh"

namespace emagls {

constexpr int MAX_SH_ORDER = 63;  // (N+1)^2 <= 4096 harmonics; reference breaks at ~85 (factorial overflow)

// Orthonormal real SH in ACN order for one direction, following the closed form that
// dependencies/Spherical-Harmonic-Transform/getSH.m:51-80 evaluates:
//   Y_nm = sqrt((2n+1)/(4pi) (n-|m|)!/(n+|m|)!) P_n^|m|(cos zen) {sqrt2 sin|m|azi, 1, sqrt2 cos m azi}
// with the Condon-Shortley sign removed.  The normalised Legendre functions come from the
// standard three-term recurrences (no factorials), evaluated from x = cos(zen) only, exactly
// like MATLAB's legendre(n, cos(zen)) (sin is taken as sqrt((1-x)(1+x)) >= 0).
// out[(n*n+n+m) * stride]; cm/sm hold cos(m*azi), sin(m*azi) for m = 0..N.
__device__ inline void real_sh_dir(int N, double x, const double* cm, const double* sm,
                                   double* out, long long stride) {
  const double s = sqrt(fmax(0.0, (1.0 - x) * (1.0 + x)));
  const double SQ2 = 1.4142135623730951;
  double pmm = 0.28209479177387814;  // sqrt(1/(4 pi))
  for (int m = 0; m <= N; ++m) {
    if (m > 0) pmm *= sqrt((2.0 * m + 1.0) / (2.0 * m)) * s;
    const double cw = (m == 0) ? 1.0 : SQ2 * cm[m];
    const double sw = (m == 0) ? 0.0 : SQ2 * sm[m];
    double p2 = 0.0, p1 = pmm;  // P(n-2,m), P(n-1,m)
    // n = m
    out[(long long)(m * m + m + m) * stride] = p1 * cw;
    if (m > 0) out[(long long)(m * m + m - m) * stride] = p1 * sw;
    for (int n = m + 1; n <= N; ++n) {
      double a = sqrt((4.0 * n * n - 1.0) / ((double)n * n - (double)m * m));
      double b = sqrt((((double)n - 1.0) * (n - 1.0) - (double)m * m) / (4.0 * (n - 1.0) * (n - 1.0) - 1.0));
      double p = a * (x * p1 - b * p2);
      p2 = p1; p1 = p;
      out[(long long)(n * n + n + m) * stride] = p * cw;
      if (m > 0) out[(long long)(n * n + n - m) * stride] = p * sw;
    }
  }
}

// Complex SH (getSH.m:25-49): Y_nm = N_nm P_n^m(x) e^{i m azi} for m >= 0 with the
// Condon-Shortley phase kept, Y_n,-m = (-1)^m conj(Y_nm).  out is interleaved complex.
__device__ inline void complex_sh_dir(int N, double x, const double* cm, const double* sm,
                                      cplx* out, long long stride) {
  const double s = sqrt(fmax(0.0, (1.0 - x) * (1.0 + x)));
  double pmm = 0.28209479177387814;
  for (int m = 0; m <= N; ++m) {
    if (m > 0) pmm *= -sqrt((2.0 * m + 1.0) / (2.0 * m)) * s;  // CS phase: (-1)^m
    double p2 = 0.0, p1 = pmm;
    const double sgn = (m & 1) ? -1.0 : 1.0;
    for (int n = m; n <= N; ++n) {
      double p;
      if (n == m) p = p1;
      else {
        double a = sqrt((4.0 * n * n - 1.0) / ((double)n * n - (double)m * m));
        double b = sqrt((((double)n - 1.0) * (n - 1.0) - (double)m * m) / (4.0 * (n - 1.0) * (n - 1.0) - 1.0));
        p = a * (x * p1 - b * p2);
        p2 = p1; p1 = p;
      }
      cplx y = mk(p * cm[m], p * sm[m]);
      out[(long long)(n * n + n + m) * stride] = y;
      if (m > 0) out[(long long)(n * n + n - m) * stride] = mk(sgn * y.x, -sgn * y.y);
    }
  }
}

// Spherical Bessel functions j_n(x), y_n(x) for n = -1..nmax (index n+1), x > 0.
// y_n: upward recurrence (stable).  j_n: upward while n < x, Miller's backward recurrence
// otherwise, normalised against the closed forms of j_0 / j_1.
// Replaces MATLAB's besselj/bessely(n+0.5, x) * sqrt(pi/(2x))
// (dependencies/Array-Response-Simulator/sph_besselj.m:14, sph_bessely.m:11).
constexpr int MAX_BESSEL = MAX_SH_ORDER + 3;
__device__ inline void sph_bessel_jy(int nmax, double x, double* j, double* y) {
  double sx, cx;
  sincos(x, &sx, &cx);
  const double ix = 1.0 / x;
  // y: y_{-1} = sin x / x, y_0 = -cos x / x
  y[0] = sx * ix;
  y[1] = -cx * ix;
  for (int n = 0; n < nmax; ++n) y[n + 2] = (2.0 * n + 1.0) * ix * y[n + 1] - y[n];
  // j
  const double j0 = sx * ix;
  const double j1 = (sx * ix - cx) * ix;
  j[0] = cx * ix;  // j_{-1}
  j[1] = j0;
  if (nmax >= 1) j[2] = j1;
  if ((double)nmax < x) {
    for (int n = 1; n < nmax; ++n) j[n + 2] = (2.0 * n + 1.0) * ix * j[n + 1] - j[n];
  } else {
    // Miller: start high enough that the trial values are 1e-300-safe and converged
    int nstart = nmax + 16 + (int)(sqrt(40.0 * (nmax + 1.0)));
    if ((double)nstart < x + 30.0) nstart = (int)x + 30;
    double fp1 = 0.0, f = 1e-280, fm1;
    double jm[MAX_BESSEL + 1];
    for (int n = nstart; n >= 1; --n) {
      fm1 = (2.0 * n + 1.0) * ix * f - fp1;   // f_{n-1}
      fp1 = f; f = fm1;
      if (n - 1 <= nmax) jm[n - 1] = f;
      if (fabs(f) > 1e250) {  // rescale to avoid overflow
        f *= 1e-250; fp1 *= 1e-250;
        for (int q = n - 1; q <= nmax; ++q) jm[q] *= 1e-250;
      }
    }
    // jm[n] proportional to j_n; normalise with the larger of j0, j1
    double scale = (fabs(j0) >= fabs(j1)) ? j0 / jm[0] : j1 / jm[1];
    for (int n = 0; n <= nmax; ++n) j[n + 1] = jm[n] * scale;
    j[1] = j0;
    if (nmax >= 1 && fabs(j0) < fabs(j1)) j[2] = j1;
  }
}

// b_n(kr) for n = 0..N following dependencies/Array-Response-Simulator/sphModalCoeffs.m:25-59
// (rigid: 4 pi i^n (j_n - (j_n'/h_n') h_n), h_n = j_n - i y_n, f_n' = (n f_{n-1} - (n+1) f_{n+1})/(2n+1);
// open: 4 pi i^n j_n), including the kr == 0 override and NaN -> 0.
__device__ inline void modal_coeffs(int N, double kr, int array_type, cplx* b) {
  if (kr == 0.0) {
    for (int n = 0; n <= N; ++n) b[n] = mk(n == 0 ? 12.566370614359172 : 0.0, 0.0);
    return;
  }
  double j[MAX_BESSEL + 1], y[MAX_BESSEL + 1];
  sph_bessel_jy(N + 1, kr, j, y);
  const double FOURPI = 12.566370614359172;
  for (int n = 0; n <= N; ++n) {
    double jn = j[n + 1], yn = y[n + 1];
    cplx val;
    if (array_type == 1) {  // open
      val = mk(jn, 0.0);
    } else {
      double inv = 1.0 / (2.0 * n + 1.0);
      double djn = inv * (n * j[n] - (n + 1.0) * j[n + 2]);
      double dyn = inv * (n * y[n] - (n + 1.0) * y[n + 2]);
      cplx hn = mk(jn, -yn), dhn = mk(djn, -dyn);
      cplx ratio = cdiv(mk(djn, 0.0), dhn);
      cplx t = cmul(ratio, hn);
      val = mk(jn - t.x, -t.y);
    }
    // times 4 pi i^n
    cplx r;
    switch (n & 3) {
      case 0: r = mk(val.x, val.y); break;
      case 1: r = mk(-val.y, val.x); break;
      case 2: r = mk(-val.x, -val.y); break;
      default: r = mk(val.y, -val.x); break;
    }
    r = cscale(r, FOURPI);
    if (isnan(r.x) || isnan(r.y)) r = mk(0.0, 0.0);
    b[n] = r;
  }
}

}  // namespace emagls
