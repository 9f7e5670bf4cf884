// FP64 tensor-core GEMM used by every dense contraction of the design path:
//   C[m][n] = epi( sum_k A(m,k) * B(n,k) )
// CTA tile 128x128x16, 8 warps (2 x 4), warp tile 64x32 = 8x4 DMMA.8x8x4 tiles, 3-stage
// cp.async (LDGSTS) ring.  A and B may be K-contiguous ([m][k]) or K-strided ([k][m]); the
// tiles are staged in shared memory in their global orientation (padding chosen so both
// fragment patterns are bank-conflict free) so no transposes are materialised in HBM.
#pragma once
#include "common.cuh"

namespace emagls {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_STAGES = 3, GM_THREADS = 256;
constexpr int GM_LDK = GM_BK + 4;    // [row][k] orientation: row stride 20 doubles
constexpr int GM_LDM = GM_BM + 4;    // [k][row] orientation: row stride 132 doubles
constexpr int GM_TILE_DOUBLES = (GM_BM * GM_LDK > GM_BK * GM_LDM) ? GM_BM * GM_LDK : GM_BK * GM_LDM;
constexpr size_t GM_SMEM_BYTES = (size_t)GM_STAGES * 2 * GM_TILE_DOUBLES * sizeof(double);

struct GemmOperand {
  const double* p;
  long long ld;      // stride (doubles) of the non-contiguous index
  int kcontig;       // 1: element (row,k) at p[row*ld + k];  0: at p[k*ld + row]
};

struct GemmShape { int M, N, K; };

// ---- epilogues -----------------------------------------------------------------------------
struct EpiStore {
  double* C; long long ldc; double alpha;
  __device__ __forceinline__ void operator()(int m, int n, double v0, double v1, int M, int N) const {
    if (m >= M) return;
    double* q = C + (long long)m * ldc + n;
    if (n + 1 < N) {
      if ((((uintptr_t)q) & 15) == 0) { *reinterpret_cast<double2*>(q) = make_double2(alpha * v0, alpha * v1); }
      else { q[0] = alpha * v0; q[1] = alpha * v1; }
    } else if (n < N) q[0] = alpha * v0;
  }
};

// MagLS phase continuation, lib/getEMagLs2Filters.m:95-103: columns (2j, 2j+1) hold (re, im)
// of y = W(k-1,:) * pwGrid for direction m and problem/ear j; the target written back is
// abs(H(k,m)) * exp(1i*angle(y)) (real part only for the Nyquist bin).
struct EpiPhase {
  double* T; long long ldt;
  const double* absH;        // [set][ear][dir] for this bin
  long long abs_set_stride;  // doubles between sets
  long long abs_ear_stride;  // doubles between ears
  int orient_per_set;
  int nyquist;
  __device__ __forceinline__ void operator()(int m, int n, double re, double im, int M, int N) const {
    if (m >= M || n >= N) return;
    int j = n >> 1;                 // problem*2 + ear
    int ear = j & 1, prob = j >> 1;
    int set = prob / orient_per_set;
    double mag = absH[(long long)set * abs_set_stride + (long long)ear * abs_ear_stride + m];
    double a2 = fma(re, re, im * im);
    double tr, ti;
    if (a2 > 0.0) {
      double inv = mag / sqrt(a2);
      tr = re * inv; ti = im * inv;
    } else { tr = mag; ti = 0.0; }  // angle(0) = 0
    if (nyquist) ti = 0.0;
    double* q = T + (long long)m * ldt + n;
    *reinterpret_cast<double2*>(q) = make_double2(tr, ti);  // n even, ldt even, T 16B aligned
  }
};

// ---- tile loader ---------------------------------------------------------------------------
// Loads a [ROWS x BK] tile of an operand into shared memory (zero-filled out of range).
template <int ROWS>
__device__ __forceinline__ void load_tile(double* s, const GemmOperand& op, int row0, int k0,
                                          int nrows, int K, bool vec16, int tid) {
  if (op.kcontig) {
    // smem [row][GM_LDK]; chunks run along k
    if (vec16) {
      constexpr int CH = GM_BK / 2;                   // 8 chunks per row
      for (int c = tid; c < ROWS * CH; c += GM_THREADS) {
        int r = c / CH, kk = (c % CH) * 2;
        int gr = row0 + r, gk = k0 + kk;
        bool ok = (gr < nrows) && (gk < K);
        const double* g = op.p + (long long)(ok ? gr : 0) * op.ld + (ok ? gk : 0);
        cp_async16(s + r * GM_LDK + kk, g, ok);
      }
    } else {
      for (int c = tid; c < ROWS * GM_BK; c += GM_THREADS) {
        int r = c / GM_BK, kk = c % GM_BK;
        int gr = row0 + r, gk = k0 + kk;
        bool ok = (gr < nrows) && (gk < K);
        const double* g = op.p + (long long)(ok ? gr : 0) * op.ld + (ok ? gk : 0);
        cp_async8(s + r * GM_LDK + kk, g, ok);
      }
    }
  } else {
    // smem [k][GM_LDM]; chunks run along row
    if (vec16) {
      constexpr int CH = ROWS / 2;
      for (int c = tid; c < GM_BK * CH; c += GM_THREADS) {
        int kk = c / CH, r = (c % CH) * 2;
        int gr = row0 + r, gk = k0 + kk;
        bool ok = (gr < nrows) && (gk < K);
        const double* g = op.p + (long long)(ok ? gk : 0) * op.ld + (ok ? gr : 0);
        cp_async16(s + kk * GM_LDM + r, g, ok);
      }
    } else {
      for (int c = tid; c < GM_BK * ROWS; c += GM_THREADS) {
        int kk = c / ROWS, r = c % ROWS;
        int gr = row0 + r, gk = k0 + kk;
        bool ok = (gr < nrows) && (gk < K);
        const double* g = op.p + (long long)(ok ? gk : 0) * op.ld + (ok ? gr : 0);
        cp_async8(s + kk * GM_LDM + r, g, ok);
      }
    }
  }
}

template <class Epi>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_f64_kernel(GemmOperand A, GemmOperand B, GemmShape sh, int vecA, int vecB, Epi epi) {
  extern __shared__ __align__(16) double gsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps
  const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
  const int ktiles = (sh.K + GM_BK - 1) / GM_BK;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  auto stageA = [&](int s) { return gsm + (size_t)s * 2 * GM_TILE_DOUBLES; };
  auto stageB = [&](int s) { return gsm + (size_t)s * 2 * GM_TILE_DOUBLES + GM_TILE_DOUBLES; };

#pragma unroll
  for (int s = 0; s < GM_STAGES - 1; ++s) {
    if (s < ktiles) {
      load_tile<GM_BM>(stageA(s), A, m0, s * GM_BK, sh.M, sh.K, vecA, tid);
      load_tile<GM_BN>(stageB(s), B, n0, s * GM_BK, sh.N, sh.K, vecB, tid);
    }
    cp_async_commit();
  }

  const int lr = lane >> 2, lk = lane & 3;
  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<GM_STAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt + STAGES - 1 into the slot freed by iteration kt - 1
      int nk = kt + GM_STAGES - 1;
      if (nk < ktiles) {
        int s = nk % GM_STAGES;
        load_tile<GM_BM>(stageA(s), A, m0, nk * GM_BK, sh.M, sh.K, vecA, tid);
        load_tile<GM_BN>(stageB(s), B, n0, nk * GM_BK, sh.N, sh.K, vecB, tid);
      }
      cp_async_commit();
    }
    const double* As = stageA(kt % GM_STAGES);
    const double* Bs = stageB(kt % GM_STAGES);
#pragma unroll
    for (int k4 = 0; k4 < GM_BK / 4; ++k4) {
      double a[8], b[4];
      const int kk = k4 * 4 + lk;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int r = wm * 64 + i * 8 + lr;
        a[i] = A.kcontig ? As[r * GM_LDK + kk] : As[kk * GM_LDM + r];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = wn * 32 + j * 8 + lr;
        b[j] = B.kcontig ? Bs[c * GM_LDK + kk] : Bs[kk * GM_LDM + c];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + wm * 64 + i * 8 + lr;
      int n = n0 + wn * 32 + j * 8 + 2 * lk;
      epi(m, n, acc[i][j][0], acc[i][j][1], sh.M, sh.N);
    }
}

// vec16 is legal when the contiguous extent, the stride and the base pointer are all even/aligned
inline int gemm_vec_ok(const GemmOperand& op, int rows, int K) {
  bool aligned = (((uintptr_t)op.p) & 15) == 0 && (op.ld % 2 == 0);
  int contig_extent = op.kcontig ? K : rows;
  return aligned && (contig_extent % 2 == 0);
}

template <class Epi>
inline cudaError_t launch_gemm(cudaStream_t st, GemmOperand A, GemmOperand B, GemmShape sh, Epi epi) {
  static bool attr_set = false;
  auto kern = gemm_f64_kernel<Epi>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((sh.N + GM_BN - 1) / GM_BN, (sh.M + GM_BM - 1) / GM_BM);
  kern<<<grid, GM_THREADS, GM_SMEM_BYTES, st>>>(A, B, sh, gemm_vec_ok(A, sh.M, sh.K), gemm_vec_ok(B, sh.N, sh.K), epi);
  return cudaGetLastError();
}

}  // namespace emagls
