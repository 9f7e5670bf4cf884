// Host orchestration of the factored eMagLS design path (getEMagLs2Filters / getEMagLsFilters).
//
// Reference flow being reproduced (lib/getEMagLs2Filters.m:44-135), restructured for a batch of
// P = num_sets * num_orient problems:
//   once per grid      : Y_h = getSH(simN, grid) = Q R                       (Householder, on device)
//   once per array     : b_n(kr_k) for all bins                              (getSMAIRMatrix.m:107)
//   once per orientation: Ym_o = L * getSH(simN, R_o^T mics);  E_o = R-blocks * Ym_o^T
//   once per HRTF set  : group delays, H = fft(h) .* ramp, |H|, H*Q for the LS bins
//   per bin k          : C = sum_n b_n(k) E_n -> TSQR -> clipped inverse   (factor kernel, SM-local)
//                        LS bins:    W_k = (H_k Q) conj(Q_C) Pb
//                        MagLS bins: y = W_{k-1} C^T Q^T;  t = |H_k| y/|y|;  W_k = (t Q) conj(Q_C) Pb
//   tail               : DC fix, ifft, sub-sample shift, crop, fade folded into one GEMM per ear
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "engine.h"
#include "gemm.cuh"
#include "special.cuh"

namespace emagls {

namespace {

struct EpiRamp {  // H = fft(h) .* exp_omega  (applySubsampleDelay.m:10-17), columns (2k, 2k+1) = (re, im)
  double* C; long long ldc; const cplx* ramp;
  __device__ __forceinline__ void operator()(int m, int n, double re, double im, int M, int N) const {
    if (m >= M || n >= N) return;
    cplx r = ramp[n >> 1];
    double* q = C + (long long)m * ldc + n;
    *reinterpret_cast<double2*>(q) = make_double2(re * r.x - im * r.y, re * r.y + im * r.x);
  }
};

double median_of(std::vector<double> v) {
  size_t n = v.size();
  std::sort(v.begin(), v.end());
  return (n & 1) ? v[n / 2] : 0.5 * (v[n / 2 - 1] + v[n / 2]);
}

}  // namespace

void design_factored(emagls_ctx* h, const emagls_config& cfg, const DesignArgs& a) {
  cudaStream_t st = h->stream;
  // ---------------- reference arithmetic on the scalar parameters (lib/getEMagLs2Filters.m:42-48)
  EM_REQUIRE(a.T > 0 && a.D > 0 && a.M > 0 && a.len > 0, "empty input");
  EM_REQUIRE(a.len >= a.T, "len too short");  // lib/getEMagLs2Filters.m:42
  EM_REQUIRE(a.len % 2 == 0, "len must be even");
  EM_REQUIRE(a.num_sets >= 1 && a.num_orient >= 1, "empty batch");
  EM_REQUIRE(a.rotations != nullptr || a.num_orient == 1, "num_orient > 1 needs rotations");
  EM_REQUIRE(cfg.basis == EMAGLS_BASIS_REAL || a.variant == Variant::EMAGLS2,
             "complex basis for SH-domain output is not built yet");
  const int nfft = std::min(cfg.nfft_max_len, 2 * a.len);
  EM_REQUIRE(nfft % 2 == 0, "nfft must be even");  // getSMAIRMatrix.m:89
  EM_REQUIRE(nfft / 2 >= a.len / 2, "len exceeds NFFT_MAX_LEN (reference indexes out of range here)");
  const int K = nfft / 2 + 1;
  const double df = (a.fs / 2.0) / (double)(K - 1);
  const double f_cut = std::max(cfg.f_cut_min, 500.0 * a.order);
  const int k_cut = (int)std::ceil(f_cut / df);  // 1-based MATLAB index
  const int simN = std::max(a.order, (int)std::ceil(a.fs * M_PI * a.mic_radius / cfg.speed_of_sound));
  EM_REQUIRE(simN <= MAX_SH_ORDER, "simulation order too high");
  const int S = (simN + 1) * (simN + 1);
  const int Mc = (a.variant == Variant::EMAGLS2) ? a.M : (a.order + 1) * (a.order + 1);
  EM_REQUIRE(Mc <= 64, "more than 64 output channels are not supported");
  if (a.D < S)
    throw Fail{EMAGLS_ERR_UNSUPPORTED, "HRIR grid has fewer directions than simulation harmonics"};
  EM_REQUIRE(S >= Mc, "fewer simulation harmonics than channels");
  const int P = a.num_sets * a.num_orient;
  const int D = a.D, T = a.T;
  // bins (0-based): LS 1 .. kls1-1, MagLS kls1 .. K-1, with kls1 = k_cut - 1
  const int kls1 = std::min(std::max(k_cut - 1, 1), K);
  const int nLS = kls1 - 1;

  Arena ar(st);
  ProfSpan* setup_span = new ProfSpan(h, EM_PROF_SETUP);
  struct SpanGuard { ProfSpan*& p; ~SpanGuard() { delete p; p = nullptr; } } setup_guard{setup_span};
  // ---------------- grid: Y_h = Q R
  double* Yh = ar.get<double>((size_t)S * D);
  double* Q = ar.get<double>((size_t)S * D);
  double* R = ar.get<double>((size_t)S * S);
  {
    double* work = ar.get<double>((size_t)S * D + 2 * S);
    EM_CUDA(launch_sh_angles(st, simN, a.grid_azi, a.grid_zen, D, 0, Yh));
    h->launches += 1;
    EM_CUDA(launch_householder_qr(st, Yh, D, S, Q, R, work, &h->launches));
  }
  // ---------------- array: b_n table (minus sign, Nyquist real: getSMAIRMatrix.m:107,115-117)
  std::vector<double> kr(K);
  for (int k = 0; k < K; ++k) {
    double f = (double)k * df;
    kr[k] = 2.0 * M_PI * f / cfg.speed_of_sound * a.mic_radius;
  }
  double* d_kr = ar.upload(kr.data(), K);
  cplx* bn = ar.get<cplx>((size_t)K * (simN + 1));
  EM_CUDA(launch_modal(st, simN, d_kr, K, cfg.array_type, -1.0, 1, bn, simN + 1, 1));
  h->launches += 1;

  // ---------------- orientations: Ym_o (and L * Ym_o for SH-domain output), E rows
  std::vector<int> rowoff(S), roword(S);
  long long Etot = 0;
  {
    int i = 0;
    for (int n = 0; n <= simN; ++n)
      for (int m = -n; m <= n; ++m, ++i) {
        roword[i] = n;
        rowoff[i] = (int)Etot;
        Etot += (long long)(simN + 1 - n) * Mc;
      }
  }
  int* d_rowoff = ar.upload(rowoff.data(), S);
  int* d_roword = ar.upload(roword.data(), S);
  double* E = ar.get<double>((size_t)a.num_orient * Etot);
  {
    double* Ym = ar.get<double>((size_t)a.num_orient * a.M * S);
    EM_CUDA(launch_sh_mics(st, simN, a.mic_azi, a.mic_zen, a.M, a.rotations, a.num_orient, Ym));
    h->launches += 1;
    const double* Yeff = Ym;
    if (a.variant == Variant::EMAGLS_SH) {
      throw Fail{EMAGLS_ERR_UNSUPPORTED, "SH-domain eMagLS is wired in a later step"};
    }
    for (int o0 = 0; o0 < a.num_orient; o0 += 32768) {
      int nb = std::min(32768, a.num_orient - o0);
      EM_CUDA(launch_build_E(st, R, S, simN, Yeff + (size_t)o0 * Mc * S, Mc, nb, d_rowoff, d_roword, Etot,
                             E + (size_t)o0 * Etot));
      h->launches += 1;
    }
  }

  // ---------------- HRTF sets: group delay, H, |H|, H*Q
  std::vector<double> grpD((size_t)a.num_sets * 2);
  {
    const int nchunk = 32;
    double* partial = ar.get<double>((size_t)nchunk * T);
    double* hsum = ar.get<double>((size_t)a.num_sets * 2 * T);
    double* gd = ar.get<double>((size_t)a.num_sets * 2 * K);
    for (int s = 0; s < a.num_sets; ++s)
      for (int e = 0; e < 2; ++e) {
        const double* hp = (e == 0 ? a.hL : a.hR) + (size_t)s * T * D;
        EM_CUDA(launch_colsum(st, hp, T, D, partial, nchunk, hsum + ((size_t)s * 2 + e) * T));
        EM_CUDA(launch_grpdelay(st, hsum + ((size_t)s * 2 + e) * T, T, K, a.fs, gd + ((size_t)s * 2 + e) * K));
        h->launches += 3;
      }
    std::vector<double> gdh((size_t)a.num_sets * 2 * K);
    EM_CUDA(cudaMemcpyAsync(gdh.data(), gd, gdh.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    EM_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < a.num_sets * 2; ++i)
      grpD[i] = median_of(std::vector<double>(gdh.begin() + (size_t)i * K, gdh.begin() + (size_t)(i + 1) * K));
  }
  double* absH = ar.get<double>((size_t)a.num_sets * 2 * K * D);           // [set][ear][K][D]
  double* Tls = ar.get<double>((size_t)a.num_sets * 2 * std::max(nLS, 1) * 2 * S);  // [set][ear][kls][c][S]
  {
    double* tw = ar.get<double>((size_t)2 * K * T);
    EM_CUDA(launch_dft_twiddle(st, K, T, nfft, tw));
    h->launches += 1;
    double* Hd = ar.get<double>((size_t)D * 2 * K);
    std::vector<cplx> ramp((size_t)a.num_sets * 2 * K);
    for (int i = 0; i < a.num_sets * 2; ++i)
      for (int k = 0; k < K; ++k) {
        double omega = (double)k * (0.5 / (double)(K - 1));
        double ang = -2.0 * M_PI * omega * (-grpD[i]);
        cplx r = mk(std::cos(ang), std::sin(ang));
        if (k == K - 1) r.y = 0.0;
        ramp[(size_t)i * K + k] = r;
      }
    cplx* d_ramp = ar.upload(ramp.data(), ramp.size());
    for (int s = 0; s < a.num_sets; ++s)
      for (int e = 0; e < 2; ++e) {
        const double* hp = (e == 0 ? a.hL : a.hR) + (size_t)s * T * D;
        GemmOperand A{hp, T, 1}, B{tw, T, 1};
        EpiRamp epi{Hd, 2LL * K, d_ramp + ((size_t)s * 2 + e) * K};
        EM_CUDA(launch_gemm(st, A, B, GemmShape{D, 2 * K, T}, epi));
        EM_CUDA(launch_abs_transpose(st, Hd, D, K, absH + ((size_t)s * 2 + e) * K * D));
        h->launches += 2;
        if (nLS > 0) {
          GemmOperand A2{Hd + 2, 2LL * K, 0}, B2{Q, D, 1};
          EpiStore st2{Tls + ((size_t)s * 2 + e) * nLS * 2 * S, S, 1.0};
          EM_CUDA(launch_gemm(st, A2, B2, GemmShape{2 * nLS, S, D}, st2));
          h->launches += 1;
        }
      }
  }

  delete setup_span; setup_span = nullptr;
  // ---------------- the hot loop over bins
  const BlockPlan bp = make_block_plan(S, Mc);
  int G = std::max(4, (1776 + P - 1) / P);
  G = std::min(G, 32);
  G = std::min(G, K - 1);
  OperatorSet ops;
  ops.v_stride = (long long)Mc * S;
  ops.tau_stride = (long long)bp.nblk * bp.MC;
  ops.rc_stride = (long long)Mc * Mc;
  ops.pb_stride = (long long)Mc * Mc;
  ops.V = ar.get<cplx>((size_t)P * G * ops.v_stride);
  ops.tau = ar.get<cplx>((size_t)P * G * ops.tau_stride);
  ops.Rc = ar.get<cplx>((size_t)P * G * ops.rc_stride);
  ops.Pb = ar.get<cplx>((size_t)P * G * ops.pb_stride);
  ops.info = ar.get<int>((size_t)P * G);
  RowSource src{};
  src.E = E; src.Etot = Etot; src.rowoff = d_rowoff; src.roword = d_roword; src.bn = bn; src.N = simN;
  // E is per orientation; problems of different HRTF sets share it: problem p -> orientation p % num_orient.
  // The factor kernel indexes E by problem, so factor per set when num_sets > 1 (operators are
  // identical across sets; only computed once and reused).
  const int PF = a.num_orient;  // problems factorised
  cplx* Wsp = a.spectra ? reinterpret_cast<cplx*>(a.spectra) : ar.get<cplx>((size_t)2 * P * Mc * K);
  const long long w_ear = (long long)P * Mc * K;
  double* Cv = ar.get<double>((size_t)4 * P * S);
  double* Tt = ar.get<double>((size_t)D * 4 * P);
  double* tq = ar.get<double>((size_t)4 * P * S);
  EM_CUDA(cudaMemsetAsync(Wsp, 0, (size_t)2 * P * Mc * K * sizeof(cplx), st));

  for (int g0 = 1; g0 < K; g0 += G) {
    const int Gn = std::min(G, K - g0);
    {
      ProfSpan ps(h, EM_PROF_FACTOR);
      EM_CUDA(launch_factor(st, bp, src, ops, PF, g0, Gn, cfg.svd_regul));
    }
    h->launches += 1;
    if (getenv("EMAGLS_DEBUG_INFO")) {
      std::vector<int> info((size_t)PF * Gn);
      EM_CUDA(cudaMemcpyAsync(info.data(), ops.info, info.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
      EM_CUDA(cudaStreamSynchronize(st));
      for (int slot = 0; slot < Gn; ++slot) {
        int nfast = 0, smax = 0; long long ssum = 0;
        for (int p = 0; p < PF; ++p) { int v = info[(size_t)p * Gn + slot]; nfast += (v == 0); smax = std::max(smax, v); ssum += v; }
        fprintf(stderr, "bin %d: fast %d/%d, sweeps max %d mean %.2f\n", g0 + slot, nfast, PF, smax, (double)ssum / PF);
      }
    }
    for (int slot = 0; slot < Gn; ++slot) {
      const int kb = g0 + slot;
      if (kb < kls1) {
        ProfSpan ps(h, EM_PROF_CHAIN_BWD);
        EM_CUDA(launch_chain_bwd(st, bp, ops, slot, Gn, Tls + (size_t)(kb - 1) * 2 * S, (long long)2 * nLS * 2 * S,
                                 (long long)nLS * 2 * S, 1, a.num_orient, PF, Wsp, w_ear, K, kb, 1, P));
        h->launches += 1;
      } else {
        {
          ProfSpan ps(h, EM_PROF_CHAIN_FWD);
          EM_CUDA(launch_chain_fwd(st, bp, ops, slot, Gn, PF, Wsp, w_ear, K, kb - 1, P, Cv));
        }
        {
          ProfSpan ps(h, EM_PROF_GEMM_FWD);
          GemmOperand A3{Q, D, 0}, B3{Cv, S, 1};
          EpiPhase ep{Tt, 4LL * P, absH + (size_t)kb * D, 2LL * K * D, (long long)K * D, a.num_orient,
                      kb == K - 1 ? 1 : 0};
          EM_CUDA(launch_gemm(st, A3, B3, GemmShape{D, 4 * P, S}, ep));
        }
        {
          ProfSpan ps(h, EM_PROF_GEMM_BWD);
          GemmOperand A4{Tt, 4LL * P, 0}, B4{Q, D, 1};
          EpiStore es{tq, S, 1.0};
          EM_CUDA(launch_gemm(st, A4, B4, GemmShape{4 * P, S, D}, es));
        }
        {
          ProfSpan ps(h, EM_PROF_CHAIN_BWD);
          EM_CUDA(launch_chain_bwd(st, bp, ops, slot, Gn, tq, 0, 0, 0, a.num_orient, PF, Wsp, w_ear, K, kb, 1, P));
        }
        h->launches += 4;
      }
    }
  }

  // ---------------- tail: one GEMM per (set, ear)
  {
    ProfSpan ps(h, EM_PROF_TAIL);
    double* twT = ar.get<double>((size_t)a.len * 2 * K);
    for (int s = 0; s < a.num_sets; ++s)
      for (int e = 0; e < 2; ++e) {
        double delay = (double)(nfft / 2) + (e == 1 ? (grpD[(size_t)s * 2 + 1] - grpD[(size_t)s * 2]) : 0.0);
        EM_CUDA(launch_tail_twiddle(st, K, nfft, a.len, delay, twT));
        const double* Wse = reinterpret_cast<const double*>(Wsp + ((size_t)e * P + (size_t)s * a.num_orient) * Mc * K);
        GemmOperand A5{Wse, 2LL * K, 1}, B5{twT, 2LL * K, 1};
        double* out = (e == 0 ? a.wL : a.wR) + (size_t)s * a.num_orient * Mc * a.len;
        EpiStore es{out, a.len, 1.0};
        EM_CUDA(launch_gemm(st, A5, B5, GemmShape{a.num_orient * Mc, a.len, 2 * K}, es));
        h->launches += 2;
      }
  }
}

}  // namespace emagls
