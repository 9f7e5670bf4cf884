// binauralDecode (dependencies/binauralDecode.m:33-64): out(:,ear) = sum_ch fftfilt(w_ear(:,ch), in(:,ch)),
// i.e. the first num_samples samples of the multichannel linear convolution, as a partitioned
// overlap-save convolution.  cuFFT performs the D2Z / Z2D transforms only; the per-bin
// multiply-accumulate over channels (the HBM-bound part) is the hand-written kernel below.
#include <cufft.h>
#include <algorithm>
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

// (Re)build the three plans of the render for this geometry; cached in the handle between calls.
void ensure_plans(emagls_ctx* h, int N, int num_ch, int chunk, long long seg_len) {
  auto& rp = h->render_plans;
  if (rp.N == N && rp.num_ch == num_ch && rp.chunk == chunk && rp.fwd && rp.inv && rp.filt) return;
  if (rp.fwd) cufftDestroy(rp.fwd);
  if (rp.inv) cufftDestroy(rp.inv);
  if (rp.filt) cufftDestroy(rp.filt);
  rp = emagls_ctx::RenderPlans{};
  const int L = N / 2, F = N / 2 + 1;
  int n[1] = {N};
  cufftHandle pf = 0, pi = 0, pw = 0;
  EM_FFT(cufftPlanMany(&pw, 1, n, nullptr, 1, N, nullptr, 1, F, CUFFT_D2Z, 2 * num_ch));
  int inembed[1] = {(int)std::min<long long>(seg_len, 1 << 30)}, onembed[1] = {F};
  EM_FFT(cufftPlanMany(&pf, 1, n, inembed, 1, L, onembed, 1, F, CUFFT_D2Z, num_ch * (chunk + 1)));
  EM_FFT(cufftPlanMany(&pi, 1, n, nullptr, 1, F, nullptr, 1, N, CUFFT_Z2D, 2 * chunk));
  EM_FFT(cufftSetStream(pw, h->stream));
  EM_FFT(cufftSetStream(pf, h->stream));
  EM_FFT(cufftSetStream(pi, h->stream));
  rp.N = N; rp.num_ch = num_ch; rp.chunk = chunk; rp.fwd = pf; rp.inv = pi; rp.filt = pw;
}

// xp[ch][L + n] = in[ch][n0 + n - L ...]: segment buffer for one chunk of blocks.  For chunk
// starting at sample s0 (multiple of L) the buffer holds samples [s0 - L, s0 + nb*L + L) per
// channel, zero outside [0, num_samples).
__global__ void stage_input_kernel(const double* __restrict__ in, long long num_samples, int num_ch,
                                   long long s0, long long seg_len, double* __restrict__ xp) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = seg_len * num_ch;
  if (idx >= total) return;
  int ch = (int)(idx / seg_len);
  long long i = idx % seg_len;
  long long n = s0 + i;
  xp[idx] = (n >= 0 && n < num_samples) ? in[(long long)ch * num_samples + n] : 0.0;
}

__global__ void pad_filters_kernel(const double* __restrict__ wL, const double* __restrict__ wR, int len,
                                   int num_ch, int N, double* __restrict__ wp) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * num_ch * N) return;
  int t = idx % N, ch = (idx / N) % num_ch, ear = idx / (N * num_ch);
  const double* w = ear ? wR : wL;
  wp[idx] = (t < len) ? w[(long long)ch * len + t] : 0.0;
}

// Y[ear][b][f] = sum_ch X[ch][b][f] * Hw[ear][ch][f].  One thread per (b, f); the channel loop is
// unrolled so 8 independent 16-byte loads are in flight per thread.
__global__ void __launch_bounds__(256)
spectral_mac_kernel(const cplx* __restrict__ X, const cplx* __restrict__ Hw, int num_ch, int nbt,
                    int nb, int F, cplx* __restrict__ Y) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nb * F) return;
  int f = (int)(idx % F);
  int b = (int)(idx / F);
  const cplx* x = X + (long long)b * F + f;
  const long long xs = (long long)nbt * F;  // channel stride of X
  const cplx* hl = Hw + f;
  const cplx* hr = Hw + (long long)num_ch * F + f;
  cplx yl = mk(0.0, 0.0), yr = mk(0.0, 0.0);
  int ch = 0;
  for (; ch + 8 <= num_ch; ch += 8) {
    cplx xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = x[(long long)(ch + u) * xs];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      cfma(yl, xv[u], hl[(long long)(ch + u) * F]);
      cfma(yr, xv[u], hr[(long long)(ch + u) * F]);
    }
  }
  for (; ch < num_ch; ++ch) {
    cplx xv = x[(long long)ch * xs];
    cfma(yl, xv, hl[(long long)ch * F]);
    cfma(yr, xv, hr[(long long)ch * F]);
  }
  Y[(long long)b * F + f] = yl;
  Y[((long long)nb + b) * F + f] = yr;
}

// out[ear][row] = yseg[ear][b][L + i] / N for full-signal sample n = s0 + b*L + i, row = n - skip
__global__ void unstage_output_kernel(const double* __restrict__ yseg, int nb, int N, int L, long long s0,
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
  out[(long long)ear * out_rows + (n - skip)] = yseg[((long long)ear * nb + b) * N + L + i] * (1.0 / (double)N);
}

}  // namespace

void binaural_decode_dev(emagls_ctx* h, const double* in, long long num_samples, int num_ch,
                         const double* wL, const double* wR, int len, int compensate_delay, double* out) {
  EM_REQUIRE(in && wL && wR && out, "null argument");
  EM_REQUIRE(num_samples > 0 && num_ch > 0 && len > 0, "empty input");
  EM_REQUIRE(!compensate_delay || len % 2 == 0, "compensateDelay needs an even filter length");
  cudaStream_t st = h->stream;
  Arena ar(st);
  int N = 2;
  while (N < 2 * len) N <<= 1;
  const int L = N / 2, F = N / 2 + 1;
  const long long skip = compensate_delay ? (len / 2 - 1) : 0;  // out(del:end,:), binauralDecode.m:55-56
  const long long out_rows = num_samples - skip;
  EM_REQUIRE(out_rows > 0, "signal shorter than the compensated delay");
  const long long nblk_total = (num_samples + L - 1) / L;
  int chunk_max = 2048;
  if (const char* e = getenv("EMAGLS_RENDER_CHUNK")) chunk_max = std::max(1, atoi(e));
  const int chunk = (int)std::min<long long>(nblk_total, chunk_max);
  // every channel holds (chunk + 1) hops of L samples (+ L at the very end) so that the overlapping
  // D2Z batch has a uniform signal distance of L
  const long long seg_len = (long long)(chunk + 1) * L;
  ensure_plans(h, N, num_ch, chunk, seg_len);
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

  double* xp = ar.get<double>((size_t)seg_len * num_ch + L);
  cplx* X = ar.get<cplx>((size_t)num_ch * (chunk + 1) * F);
  cplx* Y = ar.get<cplx>((size_t)2 * chunk * F);
  double* yseg = ar.get<double>((size_t)2 * chunk * N);
  EM_CUDA(cudaMemsetAsync(xp + seg_len * num_ch, 0, (size_t)L * sizeof(double), st));
  for (long long b0 = 0; b0 < nblk_total; b0 += chunk) {
    const long long s0 = b0 * L;
    {
      ProfSpan ps(h, EM_PROF_RENDER_STAGE);
      long long total = seg_len * num_ch;
      stage_input_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, num_samples, num_ch, s0 - L, seg_len, xp);
      EM_CUDA(cudaGetLastError());
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_FFT);
      EM_FFT(cufftExecD2Z(rp.fwd, xp, reinterpret_cast<cufftDoubleComplex*>(X)));
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_MAC);
      long long total = (long long)chunk * F;
      // Y is laid out [ear][chunk][F] for the fixed-size inverse plan (a partial last chunk
      // computes a few unused blocks)
      spectral_mac_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, Hw, num_ch, chunk + 1, chunk, F, Y);
      EM_CUDA(cudaGetLastError());
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_FFT);
      EM_FFT(cufftExecZ2D(rp.inv, reinterpret_cast<cufftDoubleComplex*>(Y), yseg));
    }
    {
      ProfSpan ps(h, EM_PROF_RENDER_STAGE);
      long long total = (long long)2 * chunk * L;
      unstage_output_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(yseg, chunk, N, L, s0, num_samples, skip,
                                                                           out_rows, out);
      EM_CUDA(cudaGetLastError());
    }
    h->launches += 3;
  }
  // stream-ordered: the arena's cudaFreeAsync calls queue behind the work above
}

void destroy_render_plans(emagls_ctx* h) {
  auto& rp = h->render_plans;
  if (rp.fwd) cufftDestroy(rp.fwd);
  if (rp.inv) cufftDestroy(rp.inv);
  if (rp.filt) cufftDestroy(rp.filt);
  rp = emagls_ctx::RenderPlans{};
}

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
