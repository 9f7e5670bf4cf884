// binauralDecode (dependencies/binauralDecode.m:33-64): out(:,ear) = sum_ch fftfilt(w_ear(:,ch), in(:,ch)),
// i.e. the first num_samples samples of the multichannel linear convolution, as a partitioned
// overlap-save convolution.  cuFFT performs the D2Z / Z2D transforms only; the per-bin
// multiply-accumulate over channels (the HBM-bound part) is the hand-written kernel below.
#include <cufft.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "engine.h"

namespace emagls {
namespace {

#define EM_FFT(expr)                                                                              \
  do {                                                                                            \
    cufftResult _r = (expr);                                                                      \
    if (_r != CUFFT_SUCCESS)                                                                      \
      throw ::emagls::Fail{EMAGLS_ERR_CUDA, std::string(#expr) + ": cufft error " + std::to_string((int)_r)}; \
  } while (0)

constexpr int MAX_EDGE = 8;  // staged boundary blocks per chunk (head + tail)

void drop_plans(emagls_ctx* h) {
  auto& rp = h->render_plans;
  if (rp.fwd) cufftDestroy(rp.fwd);
  if (rp.inv) cufftDestroy(rp.inv);
  if (rp.filt) cufftDestroy(rp.filt);
  if (rp.edge) cufftDestroy(rp.edge);
  for (auto& d : rp.direct) cufftDestroy(d.second);
  rp = emagls_ctx::RenderPlans{};
}

// (Re)build the plans of the render for this geometry; cached in the handle between calls.
void ensure_plans(emagls_ctx* h, int N, int L, int num_ch, int chunk, int nbc) {
  auto& rp = h->render_plans;
  if (rp.N == N && rp.L == L && rp.num_ch == num_ch && rp.chunk == chunk && rp.fwd && rp.inv && rp.filt) return;
  drop_plans(h);
  const int F = N / 2 + 1;
  int n[1] = {N};
  cufftHandle pf = 0, pi = 0, pw = 0;
  EM_FFT(cufftPlanMany(&pw, 1, n, nullptr, 1, N, nullptr, 1, F, CUFFT_D2Z, 2 * num_ch));
  // overlapping input batches: block j of channel ch starts at (ch * nbc + j) * L
  int inembed[1] = {N}, onembed[1] = {F};
  EM_FFT(cufftPlanMany(&pf, 1, n, inembed, 1, L, onembed, 1, F, CUFFT_D2Z, num_ch * nbc));
  EM_FFT(cufftPlanMany(&pi, 1, n, nullptr, 1, F, nullptr, 1, N, CUFFT_Z2D, 2 * chunk));
  // boundary block e of every channel: in xe[ch][e][N], out X[ch][b][F]
  cufftHandle pe = 0;
  int e_in[1] = {N};
  EM_FFT(cufftPlanMany(&pe, 1, n, e_in, 1, MAX_EDGE * N, onembed, 1, nbc * F, CUFFT_D2Z, num_ch));
  EM_FFT(cufftSetStream(pw, h->stream));
  EM_FFT(cufftSetStream(pf, h->stream));
  EM_FFT(cufftSetStream(pi, h->stream));
  EM_FFT(cufftSetStream(pe, h->stream));
  rp.N = N; rp.L = L; rp.num_ch = num_ch; rp.chunk = chunk; rp.fwd = pf; rp.inv = pi; rp.filt = pw; rp.edge = pe;
}

// interior blocks of one channel straight from the caller's signal: `batch` overlapping transforms at
// distance L (cuFFT plans have a fixed batch: one plan per distinct count, normally one or two)
cufftHandle direct_plan(emagls_ctx* h, int N, int L, int batch) {
  auto& rp = h->render_plans;
  for (auto& d : rp.direct)
    if (d.first == batch) return d.second;
  const int F = N / 2 + 1;
  int n[1] = {N}, inembed[1] = {N}, onembed[1] = {F};
  cufftHandle p = 0;
  EM_FFT(cufftPlanMany(&p, 1, n, inembed, 1, L, onembed, 1, F, CUFFT_D2Z, batch));
  EM_FFT(cufftSetStream(p, h->stream));
  if (rp.direct.size() >= 4) { cufftDestroy(rp.direct.front().second); rp.direct.erase(rp.direct.begin()); }
  rp.direct.emplace_back(batch, p);
  return p;
}

// Segment buffer of one chunk of blocks: xp[ch][i] = in[ch][s_first + i] (zero outside the signal),
// i in [0, seg_len).  Block j of a channel is the N samples starting at xp[ch][j * L].
__global__ void stage_input_kernel(const double* __restrict__ in, long long num_samples, int num_ch,
                                   long long s_first, long long seg_len, long long ch_stride,
                                   double* __restrict__ xp) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = seg_len * num_ch;
  if (idx >= total) return;
  int ch = (int)(idx / seg_len);
  long long i = idx % seg_len;
  long long n = s_first + i;
  xp[(long long)ch * ch_stride + i] = (n >= 0 && n < num_samples) ? in[(long long)ch * num_samples + n] : 0.0;
}

__global__ void pad_filters_kernel(const double* __restrict__ wL, const double* __restrict__ wR, int len,
                                   int num_ch, int N, double* __restrict__ wp) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * num_ch * N) return;
  int t = idx % N, ch = (idx / N) % num_ch, ear = idx / (N * num_ch);
  const double* w = ear ? wR : wL;
  wp[idx] = (t < len) ? w[(long long)ch * len + t] : 0.0;
}

// Y[ear][b][f] = sum_ch X[ch][b][f] * Hw[ear][ch][f].  One thread per (f, BT consecutive blocks):
// each pair of filter-spectrum values (L2-resident) is used for BT blocks, and the channel loop is
// unrolled by CU so CU * (BT + 2) independent 16-byte loads are in flight per thread.
template <int BT, int CU>
__global__ void __launch_bounds__(128)
spectral_mac_kernel(const cplx* __restrict__ X, const cplx* __restrict__ Hw, int num_ch, int nbc,
                    int nb, int F, cplx* __restrict__ Y) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * BT;
  if (f >= F) return;
  const long long xs = (long long)nbc * F;  // channel stride of X
  const cplx* x = X + (long long)b0 * F + f;
  const cplx* hl = Hw + f;
  const cplx* hr = Hw + (long long)num_ch * F + f;
  cplx yl[BT], yr[BT];
#pragma unroll
  for (int t = 0; t < BT; ++t) { yl[t] = mk(0.0, 0.0); yr[t] = mk(0.0, 0.0); }
  // blocks past nb inside the same channel segment still exist (nbc > nb), so full tiles never
  // read out of range; the stores are guarded
  const int bt = min(BT, nbc - b0);
  int ch = 0;
  if (bt == BT) {
    for (; ch + CU <= num_ch; ch += CU) {
      cplx xv[CU][BT], a[CU], c[CU];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        a[u] = hl[(long long)(ch + u) * F];
        c[u] = hr[(long long)(ch + u) * F];
#pragma unroll
        for (int t = 0; t < BT; ++t) xv[u][t] = x[(long long)(ch + u) * xs + (long long)t * F];
      }
#pragma unroll
      for (int u = 0; u < CU; ++u)
#pragma unroll
        for (int t = 0; t < BT; ++t) { cfma(yl[t], xv[u][t], a[u]); cfma(yr[t], xv[u][t], c[u]); }
    }
  }
  for (; ch < num_ch; ++ch) {
    const cplx a = hl[(long long)ch * F], c = hr[(long long)ch * F];
#pragma unroll
    for (int t = 0; t < BT; ++t)
      if (t < bt) {
        const cplx xv = x[(long long)ch * xs + (long long)t * F];
        cfma(yl[t], xv, a); cfma(yr[t], xv, c);
      }
  }
#pragma unroll
  for (int t = 0; t < BT; ++t)
    if (b0 + t < nb) {
      Y[(long long)(b0 + t) * F + f] = yl[t];
      Y[((long long)nb + b0 + t) * F + f] = yr[t];
    }
}

// out[ear][row] = yseg[ear][b][ov + i] / N for full-signal sample n = s0 + b*L + i, row = n - skip
__global__ void unstage_output_kernel(const double* __restrict__ yseg, int nb, int N, int L, int ov, long long s0,
                                      long long num_samples, long long skip, long long out_rows,
                                      double* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)2 * nb * L;
  if (idx >= total) return;
  int i = (int)(idx % L);
  int b = (int)((idx / L) % nb);
  int ear = (int)(idx / ((long long)L * nb));
  long long n = s0 + (long long)b * L + i;
  if (n >= num_samples || n < skip) return;
  out[(long long)ear * out_rows + (n - skip)] = yseg[((long long)ear * nb + b) * N + ov + i] * (1.0 / (double)N);
}

// ---------------------------------------------------------------------------------------------
// Fused overlap-save block: one CTA per block of L output frames.  For every channel the N-point real
// FFT (as an N/2-point complex Stockham FFT, radix 4 with a leading radix-2 stage when log2(N/2) is odd,
// ping-pong in shared memory), the multiply-accumulate with both ears' filter spectra into two
// accumulators in shared memory, then the two inverse transforms and the store of the valid L frames.
// The channel spectra never reach HBM: the kernel reads the input once (plus the N/L overlap) and the
// L2-resident filter spectra, and writes the two output channels.
//   rfft:  X[k] = (Z[k] + conj(Z[M-k])) / 2 - (i/2) W_N^k (Z[k] - conj(Z[M-k])),  z[n] = x[2n] + i x[2n+1]
//   irfft: Zy[k] = (Y[k] + conj(Y[M-k])) + i conj(W_N^k) (Y[k] - conj(Y[M-k])),  y[2n] + i y[2n+1] = IFFT_M(Zy) / N
// WM[q] = exp(-2 pi i q / M) (q < M/2), WN[k] = exp(-2 pi i k / N) (k <= M), M = N / 2.
// ---------------------------------------------------------------------------------------------
__global__ void render_twiddle_kernel(int M, cplx* __restrict__ WM, cplx* __restrict__ WN) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M / 2) { double s_, c_; sincospi(-2.0 * (double)i / (double)M, &s_, &c_); WM[i] = mk(c_, s_); }
  if (i <= M) { double s_, c_; sincospi(-(double)i / (double)M, &s_, &c_); WN[i] = mk(c_, s_); }
}

// Stockham FFT of length M on ping-pong buffers, NB independent transforms side by side (buffer pair q at
// src + q * stride, dst + q * stride; the twiddles are shared); returns the offset-0 buffer holding the results
template <bool INV, int NB>
__device__ __forceinline__ cplx* stockham_fft(cplx* src, cplx* dst, long long stride, int M, int logM,
                                              const cplx* __restrict__ WM, int tid, int nt) {
  int p = 1;
  if (logM & 1) {
    for (int j = tid; j < M / 2; j += nt) {
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const cplx u0 = src[q * stride + j], u1 = src[q * stride + j + M / 2];
        dst[q * stride + 2 * j] = cadd(u0, u1); dst[q * stride + 2 * j + 1] = csub(u0, u1);
      }
    }
    __syncthreads();
    cplx* t_ = src; src = dst; dst = t_;
    p = 2;
  }
  const int T = M / 4;
  while (p < M) {
    const int tw = M / (4 * p);
    for (int t = tid; t < T; t += nt) {
      const int k = t & (p - 1);
      cplx w1 = WM[k * tw], w2 = WM[2 * k * tw];
      if (INV) { w1.y = -w1.y; w2.y = -w2.y; }
      const cplx w3 = cmul(w1, w2);
      const int j = ((t - k) << 2) + k;
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const cplx* sq = src + q * stride;
        cplx* dq = dst + q * stride;
        const cplx u0 = sq[t], u1 = cmul(sq[t + T], w1), u2 = cmul(sq[t + 2 * T], w2), u3 = cmul(sq[t + 3 * T], w3);
        const cplx v0 = cadd(u0, u2), v1 = csub(u0, u2), v2 = cadd(u1, u3);
        cplx v3 = csub(u1, u3);
        v3 = INV ? mk(-v3.y, v3.x) : mk(v3.y, -v3.x);          // * (+i) inverse, * (-i) forward
        dq[j] = cadd(v0, v2); dq[j + p] = cadd(v1, v3); dq[j + 2 * p] = csub(v0, v2); dq[j + 3 * p] = csub(v1, v3);
      }
    }
    __syncthreads();
    cplx* t_ = src; src = dst; dst = t_;
    p <<= 2;
  }
  return src;
}

// input block of one channel -> buffer (pairs x[2n], x[2n+1] as one complex value), asynchronously with zero fill
// outside the signal (both samples of a pair are inside or outside: the block origin and num_samples are even)
__device__ __forceinline__ void fused_load_async(cplx* buf, const double* __restrict__ x, long long s_first,
                                                 long long num_samples, int M, int tid, int nt) {
  for (int n = tid; n < M; n += nt) {
    const long long i0 = s_first + 2LL * n;
    const bool in_range = (i0 >= 0 && i0 < num_samples);
    cp_async16(buf + n, x + (in_range ? i0 : 0), in_range);
  }
}

// Two channels per pass (NB = 2): buffers [A0 | A1 | B0 | B1 | accL | accR].
__global__ void __launch_bounds__(512, 1)
fused_render_kernel(const double* __restrict__ in, long long num_samples, int num_ch, const cplx* __restrict__ Hw,
                    const cplx* __restrict__ WM, const cplx* __restrict__ WN, int N, int logM, int L, int ov,
                    long long skip, long long out_rows, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char fr_raw[];
  const int M = N / 2, F = M + 1, tid = threadIdx.x, nt = blockDim.x;
  cplx* A = reinterpret_cast<cplx*>(fr_raw);        // A0, A1
  cplx* Bf = A + 2 * M;                             // B0, B1
  cplx* accL = Bf + 2 * M;
  cplx* accR = accL + F;
  const long long b = blockIdx.x;
  const long long s_first = b * (long long)L - ov;          // even: L and ov are even
  for (int k = tid; k < F; k += nt) { accL[k] = mk(0.0, 0.0); accR[k] = mk(0.0, 0.0); }
  // ping-pong roles: the transform of the pair in `cur` uses `oth` as scratch and leaves the spectra in one of
  // the two; the other one is free while the spectra are unpacked and accumulated, so the next pair of channels
  // is prefetched into it (cp.async) and becomes `cur` of the next iteration
  cplx* cur = A;
  cplx* oth = Bf;
  fused_load_async(cur, in, s_first, num_samples, M, tid, nt);
  if (num_ch > 1) fused_load_async(cur + M, in + num_samples, s_first, num_samples, M, tid, nt);
  cp_async_commit();
  for (int ch = 0; ch < num_ch; ch += 2) {
    const int nb2 = (ch + 1 < num_ch) ? 2 : 1;
    cp_async_wait<0>();
    __syncthreads();
    const cplx* Z = (nb2 == 2) ? stockham_fft<false, 2>(cur, oth, M, M, logM, WM, tid, nt)
                               : stockham_fft<false, 1>(cur, oth, M, M, logM, WM, tid, nt);
    cplx* freeb = (Z == cur) ? oth : cur;
    if (ch + 2 < num_ch) {
      fused_load_async(freeb, in + (long long)(ch + 2) * num_samples, s_first, num_samples, M, tid, nt);
      if (ch + 3 < num_ch) fused_load_async(freeb + M, in + (long long)(ch + 3) * num_samples, s_first, num_samples, M, tid, nt);
      cp_async_commit();
    }
    for (int q = 0; q < nb2; ++q) {
      const cplx* Zq = Z + (long long)q * M;
      const cplx* hl = Hw + (long long)(ch + q) * F;
      const cplx* hr = Hw + ((long long)num_ch + ch + q) * F;
      for (int k = tid; k < F; k += nt) {
        const cplx zk = Zq[k & (M - 1)], zm = cconj(Zq[(M - k) & (M - 1)]);
        const cplx sm_ = cadd(zk, zm), df = cmul(WN[k], csub(zk, zm));
        const cplx X = mk(0.5 * (sm_.x + df.y), 0.5 * (sm_.y - df.x));     // sm/2 - (i/2) df
        cfma(accL[k], X, hl[k]);
        cfma(accR[k], X, hr[k]);
      }
    }
    oth = (freeb == cur) ? oth : cur;
    cur = freeb;
    // the barrier at the top of the next iteration orders the reads of Z before the next transform's writes
  }
  __syncthreads();
  // both ears side by side
  const double inv_n = 1.0 / (double)N;
  for (int k = tid; k < M; k += nt) {
#pragma unroll
    for (int ear = 0; ear < 2; ++ear) {
      const cplx* acc = ear ? accR : accL;
      const cplx yk = acc[k], ym = cconj(acc[M - k]);
      const cplx sm_ = cadd(yk, ym), df = cmul(cconj(WN[k]), csub(yk, ym));
      A[ear * M + k] = mk(sm_.x - df.y, sm_.y + df.x);               // sm + i df
    }
  }
  __syncthreads();
  const cplx* zy = stockham_fft<true, 2>(A, Bf, M, M, logM, WM, tid, nt);
  for (int n = tid; n < M; n += nt) {
    const int i = 2 * n;
    if (i < ov) continue;
    const long long s0 = b * (long long)L + (i - ov);
#pragma unroll
    for (int ear = 0; ear < 2; ++ear) {
      const cplx v = zy[ear * M + n];
      if (s0 < num_samples && s0 >= skip) out[(long long)ear * out_rows + (s0 - skip)] = v.x * inv_n;
      if (s0 + 1 < num_samples && s0 + 1 >= skip) out[(long long)ear * out_rows + (s0 + 1 - skip)] = v.y * inv_n;
    }
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

// Overlap-save geometry: FFT size N (power of two), overlap ov >= len - 1 (even), hop L = N - ov.
// Block b takes the N input samples starting at b*L - ov and yields output samples [b*L, (b+1)*L).
// Large N amortises the overlap (N = 8 len: 12.5 % redundant input, 0.57 spectrum bins per output
// frame instead of 1.0 at N = 2 len); the chunk (blocks per cuFFT batch) is sized so that the
// staged input, its spectra and the ear spectra of one chunk stay L2-resident between the kernels.
void binaural_decode_dev(emagls_ctx* h, const double* in, long long num_samples, int num_ch,
                         const double* wL, const double* wR, int len, int compensate_delay, double* out) {
  EM_REQUIRE(in && wL && wR && out, "null argument");
  EM_REQUIRE(num_samples > 0 && num_ch > 0 && len > 0, "empty input");
  EM_REQUIRE(!compensate_delay || len % 2 == 0, "compensateDelay needs an even filter length");
  cudaStream_t st = h->stream;
  Arena ar(st);
  const int ov = (len + 1) & ~1;
  int Nmin = 16;
  while (Nmin < 2 * len) Nmin <<= 1;
  int N = Nmin;
  {
    int want = env_int("EMAGLS_RENDER_FFT", 0);
    if (want <= 0) want = 8 * Nmin / 2;                       // default N = 8 len (next power of two)
    // no point in blocks much longer than the signal
    while (N < want && (long long)(N - ov) < num_samples) N <<= 1;
  }
  const int L = N - ov, F = N / 2 + 1;
  const long long skip = compensate_delay ? (len / 2 - 1) : 0;  // out(del:end,:), binauralDecode.m:55-56
  const long long out_rows = num_samples - skip;
  EM_REQUIRE(out_rows > 0, "signal shorter than the compensated delay");
  const long long nblk_total = (num_samples + L - 1) / L;
  const int nex = (ov + L - 1) / L;   // extra hops per channel so that the overlapping batch has a uniform distance L
  // blocks per chunk: as many as a memory budget allows (measured on B200: launch-bound below a few
  // hundred blocks; one chunk for a 10-minute signal is fastest, the intermediates do not need to stay
  // in L2).  EMAGLS_RENDER_CHUNK / EMAGLS_RENDER_WS_MB override.
  int chunk_max;
  if (const char* e = getenv("EMAGLS_RENDER_CHUNK")) {
    chunk_max = std::max(1, atoi(e));
  } else {
    size_t free_b = 0, total_b = 0;
    EM_CUDA(cudaMemGetInfo(&free_b, &total_b));
    double budget = std::min(0.4 * (double)free_b, 16.0 * 1073741824.0);
    const int ws_mb = env_int("EMAGLS_RENDER_WS_MB", 0);
    if (ws_mb > 0) budget = (double)ws_mb * 1048576.0;
    const double per_block = (double)num_ch * ((double)L * 8 + (double)F * 16) + 2.0 * F * 16 + 2.0 * N * 8;
    chunk_max = (int)std::max(1.0, std::min(65535.0 * 4, budget / per_block));
  }
  const int chunk = (int)std::min<long long>({nblk_total, (long long)chunk_max, 65535LL * 4});
  const int nbc = chunk + nex;
  const long long seg_len = (long long)nbc * L;
  ensure_plans(h, N, L, num_ch, chunk, nbc);
  const auto& rp = h->render_plans;

  // filter spectra Hw[ear][ch][F]
  cplx* Hw = ar.get<cplx>((size_t)2 * num_ch * F);
  {
    double* wp = ar.get<double>((size_t)2 * num_ch * N);
    int total = 2 * num_ch * N;
    pad_filters_kernel<<<(total + 255) / 256, 256, 0, st>>>(wL, wR, len, num_ch, N, wp);
    EM_CUDA(cudaGetLastError());
    EM_FFT(cufftExecD2Z(rp.filt, wp, reinterpret_cast<cufftDoubleComplex*>(Hw)));
    h->launches += 1;
  }

  // Fused route (EMAGLS_RENDER_FUSED=1; off by default): one kernel per call, the channel spectra stay in shared
  // memory.  Needs the FFT and both accumulators in one CTA's shared memory (N <= 4096) and 16-byte aligned
  // channels.  Measured on B200 (10 minutes, 32 channels, 512 taps; tools/gpu_render_ab.py,
  // profiles/r01_v25_render_ab.txt): identical output to 8e-16, but 11.8 ms against 9.6 ms for the cuFFT route
  // below (15.3 ms before two channels shared a pass and the next pair was prefetched with cp.async) -- the radix-4
  // shared-memory FFT makes six barrier-separated passes per channel pair with one 512-thread CTA per SM; it needs
  // register-resident radix-8/16 passes before it can win.
  if (env_int("EMAGLS_RENDER_FUSED", 0) != 0 && N <= 4096 && N >= 8 && (num_samples % 2 == 0) && (ov % 2 == 0) &&
      (reinterpret_cast<uintptr_t>(in) % 16 == 0)) {
    const int M2 = N / 2;
    int logM = 0;
    while ((1 << logM) < M2) ++logM;
    cplx* WM = ar.get<cplx>((size_t)M2 / 2 + 1);
    cplx* WN = ar.get<cplx>((size_t)M2 + 1);
    render_twiddle_kernel<<<(M2 + 1 + 255) / 256, 256, 0, st>>>(M2, WM, WN);
    EM_CUDA(cudaGetLastError());
    const size_t smem = ((size_t)4 * M2 + 2 * (M2 + 1)) * sizeof(cplx);   // two channels side by side
    static size_t set_to = 0;
    if (smem > 48 * 1024 && smem > set_to) {
      EM_CUDA(cudaFuncSetAttribute(fused_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_to = smem;
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_MAC);
      fused_render_kernel<<<(unsigned)nblk_total, 512, smem, st>>>(in, num_samples, num_ch, Hw, WM, WN, N, logM, L, ov, skip,
                                                                  out_rows, out);
      EM_CUDA(cudaGetLastError());
    }
    h->launches += 2;
    return;
  }

  // Forward transforms.  Direct route (default when every channel of the caller's signal is 16-byte
  // aligned): interior blocks are transformed in place from `in`, one cuFFT call per channel, and only
  // the blocks that overlap the ends of the signal are staged zero-padded.  Staged route: the whole
  // chunk is copied into a segment buffer first (one cuFFT call per chunk).
  const bool direct = env_int("EMAGLS_RENDER_DIRECT", 1) != 0 && (num_samples % 2 == 0) &&
                      (reinterpret_cast<uintptr_t>(in) % 16 == 0);
  const long long b_int_lo = nex;                                                   // input starts at >= 0
  const long long b_int_hi = (num_samples - N + ov >= 0) ? (num_samples - N + ov) / L : -1;  // ends inside
  // the last nex blocks of the last channel read up to N samples past the staged segments
  double* xp = direct ? ar.get<double>((size_t)MAX_EDGE * N * num_ch) : ar.get<double>((size_t)seg_len * num_ch + N);
  cplx* X = ar.get<cplx>((size_t)num_ch * nbc * F);
  cplx* Y = ar.get<cplx>((size_t)2 * chunk * F);
  double* yseg = ar.get<double>((size_t)2 * chunk * N);
  if (!direct) EM_CUDA(cudaMemsetAsync(xp + seg_len * num_ch, 0, (size_t)N * sizeof(double), st));
  constexpr int BT = 4, CU = 4;
  for (long long b0 = 0; b0 < nblk_total; b0 += chunk) {
    const long long s0 = b0 * L;
    if (direct) {
      const long long nbk = std::min<long long>(chunk, nblk_total - b0);
      const long long lo = std::max(b0, b_int_lo), hi = std::min(b0 + nbk - 1, b_int_hi);
      if (hi >= lo) {
        ProfSpan ps(h, EM_PROF_RENDER_FFT);
        cufftHandle pd = direct_plan(h, N, L, (int)(hi - lo + 1));
        for (int ch = 0; ch < num_ch; ++ch)
          EM_FFT(cufftExecD2Z(pd, const_cast<double*>(in) + (long long)ch * num_samples + lo * L - ov,
                              reinterpret_cast<cufftDoubleComplex*>(X + ((long long)ch * nbc + (lo - b0)) * F)));
      }
      // boundary blocks of this chunk, MAX_EDGE at a time
      long long eb[MAX_EDGE];
      int ne = 0;
      auto flush = [&]() {
        for (int e = 0; e < ne; ++e) {
          ProfSpan ps(h, EM_PROF_RENDER_STAGE);
          long long total = (long long)N * num_ch;
          stage_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, num_samples, num_ch, eb[e] * L - ov, N,
                                                                            (long long)MAX_EDGE * N, xp + (long long)e * N);
          EM_CUDA(cudaGetLastError());
        }
        for (int e = 0; e < ne; ++e) {
          ProfSpan ps(h, EM_PROF_RENDER_FFT);
          EM_FFT(cufftExecD2Z(rp.edge, xp + (long long)e * N, reinterpret_cast<cufftDoubleComplex*>(X + (eb[e] - b0) * F)));
        }
        h->launches += ne;
        ne = 0;
      };
      for (long long b = b0; b < b0 + nbk; ++b) {
        if (b >= lo && b <= hi) { b = hi; continue; }
        eb[ne++] = b;
        if (ne == MAX_EDGE) flush();
      }
      flush();
    } else {
      {
        ProfSpan ps(h, EM_PROF_RENDER_STAGE);
        long long total = seg_len * num_ch;
        stage_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, num_samples, num_ch, s0 - ov, seg_len,
                                                                          seg_len, xp);
        EM_CUDA(cudaGetLastError());
      }
      {
        ProfSpan ps(h, EM_PROF_RENDER_FFT);
        EM_FFT(cufftExecD2Z(rp.fwd, xp, reinterpret_cast<cufftDoubleComplex*>(X)));
      }
      h->launches += 1;
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_MAC);
      // Y is laid out [ear][chunk][F] for the fixed-size inverse plan (a partial last chunk
      // computes a few unused blocks)
      dim3 grid((unsigned)((F + 127) / 128), (unsigned)((chunk + BT - 1) / BT));
      spectral_mac_kernel<BT, CU><<<grid, 128, 0, st>>>(X, Hw, num_ch, nbc, chunk, F, Y);
      EM_CUDA(cudaGetLastError());
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_FFT);
      EM_FFT(cufftExecZ2D(rp.inv, reinterpret_cast<cufftDoubleComplex*>(Y), yseg));
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_STAGE);
      long long total = (long long)2 * chunk * L;
      unstage_output_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(yseg, chunk, N, L, ov, s0, num_samples, skip,
                                                                           out_rows, out);
      EM_CUDA(cudaGetLastError());
    }
    h->launches += 2;
  }
  // stream-ordered: the arena's cudaFreeAsync calls queue behind the work above
}

void destroy_render_plans(emagls_ctx* h) { drop_plans(h); }

}  // namespace emagls

using namespace emagls;

extern "C" {

int emagls_binaural_decode_dev(emagls_handle h, const double* in, long long num_samples, int num_ch,
                               const double* wL, const double* wR, int len, int compensate_delay,
                               double* out) {
  return guarded(h, [&] { binaural_decode_dev(h, in, num_samples, num_ch, wL, wR, len, compensate_delay, out); });
}

int emagls_binaural_decode(emagls_handle h, const double* in, long long num_samples, int num_ch,
                           const double* wL, const double* wR, int len, int compensate_delay, double* out) {
  return guarded(h, [&] {
    EM_REQUIRE(in && wL && wR && out && num_samples > 0 && num_ch > 0 && len > 0, "bad argument");
    cudaStream_t st = h->stream;
    Arena ar(st);
    const long long skip = compensate_delay ? (len / 2 - 1) : 0;
    const long long out_rows = num_samples - skip;
    EM_REQUIRE(out_rows > 0, "signal shorter than the compensated delay");
    double* d_in = ar.upload(in, (size_t)num_samples * num_ch);
    double* d_wL = ar.upload(wL, (size_t)len * num_ch);
    double* d_wR = ar.upload(wR, (size_t)len * num_ch);
    double* d_out = ar.get<double>((size_t)out_rows * 2);
    binaural_decode_dev(h, d_in, num_samples, num_ch, d_wL, d_wR, len, compensate_delay, d_out);
    EM_CUDA(cudaMemcpyAsync(out, d_out, (size_t)out_rows * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    EM_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
