// Timing + digit checksum of the forward direction-grid product of the MagLS recursion (launch_oz_fwd:
// int8 tensor-core GEMM with the fused phase continuation / digit slicing epilogue) on synthetic digits of the
// BASELINE config-2 shape, one epilogue variant per process (EMAGLS_OZ_FWD=scaled|raw|<unset>).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/microbench/bin/oz_fwd_bench
//        tools/microbench/oz_fwd_bench.cu -lcuda ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../../emagls_b200/csrc/ozaki_kernels.cu"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void unused_fnv_kernel(const uint8_t* p, size_t n, unsigned long long* out) {
  // order-independent checksum: sum of byte * (index hash)
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += (unsigned long long)p[i] * ((i * 0x9E3779B97F4A7C15ull) >> 17 | 1ull);
  atomicAdd(out, acc);
}

int main(int argc, char** argv) {
  const int P = argc > 1 ? atoi(argv[1]) : 3600, sets = argc > 2 ? atoi(argv[2]) : 1, reps = argc > 3 ? atoi(argv[3]) : 20;
  const int long_mode = argc > 5 ? atoi(argv[5]) : -1;   // >= 0: run this variant for about five seconds (clock sampling)
  const int T = argc > 4 ? atoi(argv[4]) : 6;
  const int D = 2702, S = 400, K = 8;
  const int rows = 4 * P * sets, KpS = emagls::oz_pad32(S), KpD = emagls::oz_pad32(D);
  std::mt19937_64 rng(7);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> hY((size_t)D * S), hC((size_t)rows * S), hAbs((size_t)sets * 2 * K * D);
  for (auto& x : hY) x = nd(rng);
  for (size_t i = 0; i < hC.size(); ++i) hC[i] = nd(rng) * std::pow(2.0, (double)((i / S) % 9) - 4.0);
  for (auto& x : hAbs) x = std::fabs(nd(rng)) + 0.01;
  double *dY, *dC, *dAbs, *sY, *sC, *sT, *up, *sc;
  int8_t *qY, *qC, *qT;
  CK(cudaMalloc(&dY, hY.size() * 8)); CK(cudaMalloc(&dC, hC.size() * 8)); CK(cudaMalloc(&dAbs, hAbs.size() * 8));
  CK(cudaMalloc(&sY, D * 8)); CK(cudaMalloc(&sC, (size_t)rows * 8)); CK(cudaMalloc(&sT, (size_t)rows * 8));
  CK(cudaMalloc(&up, (size_t)sets * 2 * K * 8)); CK(cudaMalloc(&sc, (size_t)sets * 2 * K * 8));
  CK(cudaMalloc(&qY, (size_t)T * D * KpS)); CK(cudaMalloc(&qC, (size_t)T * rows * KpS)); CK(cudaMalloc(&qT, (size_t)T * rows * KpD));
  CK(cudaMemcpy(dY, hY.data(), hY.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAbs, hAbs.data(), hAbs.size() * 8, cudaMemcpyHostToDevice));
  CK(emagls::launch_slice_rows(0, dY, S, 1, D, S, KpS, T, qY, sY));
  CK(emagls::launch_slice_rows(0, dC, S, 1, rows, S, KpS, T, qC, sC));
  CK(emagls::launch_row_scale(0, dAbs, (long long)sets * 2 * K, D, up, sc));
  CK(cudaMemset(qT, 0, (size_t)T * rows * KpD));
  const int kb = 3;
  emagls::OzFwdArgs fa{qY, sY, D, KpS, qC, sC, rows, T, qT, KpD, sT, dAbs + (size_t)kb * D, 2LL * K * D, (long long)K * D,
                       up + kb, sc + kb, K, P, 0};
  using namespace emagls;
  using Ring2 = oz::TileCfg<oz::TILE_N, 2>;
  auto run = [&](int mode) -> cudaError_t {
    if (T == 6) {
      if (mode == 0) return oz_fwd_t<6, EpiPhaseSlice<6>>(0, fa);
      if (mode == 1) return oz_fwd_t<6, EpiPhaseSliceRaw<6>>(0, fa);
      if (mode == 3) return oz_fwd_t<6, EpiPhaseSliceRaw<6>, oz::TileWide>(0, fa);
      if (mode == 4) return oz_fwd_t<6, EpiPhaseSliceRaw<6>, oz::TileCfg<48, 3>>(0, fa);
      return oz_fwd_t<6, EpiPhaseSliceTma<6>, Ring2>(0, fa);
    }
    if (mode == 0) return oz_fwd_t<4, EpiPhaseSlice<4>>(0, fa);
    if (mode == 1) return oz_fwd_t<4, EpiPhaseSliceRaw<4>>(0, fa);
    if (mode == 3) return oz_fwd_t<4, EpiPhaseSliceRaw<4>, oz::TileWide>(0, fa);
    if (mode == 4) return oz_fwd_t<4, EpiPhaseSliceRaw<4>, oz::TileCfg<48, 3>>(0, fa);
    return oz_fwd_t<4, EpiPhaseSliceTma<4>, Ring2>(0, fa);
  };
  const char* names[5] = {"scaled", "raw", "tma", "raw80", "raw48"};
  const size_t nq = (size_t)T * rows * KpD;
  std::vector<int8_t> ref(nq), got(nq);
  std::vector<double> sref(rows), sgot(rows);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  if (long_mode >= 0) {
    for (int i = 0; i < 12000; ++i) CK(run(long_mode));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 2000; ++i) CK(run(long_mode));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float msl = 0;
    CK(cudaEventElapsedTime(&msl, e0, e1));
    printf("long run %s: %.4f ms per launch after 12000 launches\n", names[long_mode], msl / 2000);
    return 0;
  }
  for (int mode = 0; mode < 5; ++mode) {
    CK(cudaMemset(qT, 0, nq));
    CK(cudaMemset(sT, 0, (size_t)rows * 8));
    CK(run(mode));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), qT, nq, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sgot.data(), sT, (size_t)rows * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0, bad_s = 0, shown = 0;
    double maxdiff = 0.0;
    if (mode == 0) { ref = got; sref = sgot; }
    else {
      for (int r = 0; r < rows; ++r) {
        bad_s += (sgot[r] != sref[r]);
        for (int m = 0; m < KpD; ++m) {
          double va = 0, vb = 0; bool neq = false;
          for (int s = 0; s < T; ++s) {
            const size_t i = ((size_t)s * rows + r) * KpD + m;
            va += ref[i] * std::pow(256.0, -s); vb += got[i] * std::pow(256.0, -s);
            neq |= ref[i] != got[i];
          }
          if (neq) {
            ++bad; maxdiff = std::max(maxdiff, std::fabs(va - vb));
            if (shown < 6) { printf("   mismatch row %d col %d: ref %.17g got %.17g\n", r, m, va, vb); ++shown; }
          }
        }
      }
    }
    for (int dbg : {1, 2, 3, 4, 6}) {   // decomposition: 1 no epilogue, 2 no MMAs, 3 operand ring only, 4 drain only, 6 drain only without MMAs
      oz_fwd_debug = dbg;
      CK(run(mode));
      CK(cudaEventRecord(e0));
      for (int i = 0; i < reps; ++i) CK(run(mode));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float msd = 0;
      CK(cudaEventElapsedTime(&msd, e0, e1));
      printf("   %-6s dbg=%d: %.4f ms\n", names[mode], dbg, msd / reps);
    }
    oz_fwd_debug = 0;
    for (int i = 0; i < 3; ++i) CK(run(mode));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) CK(run(mode));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double ops = 2.0 * D * (double)rows * KpS * (T * (T + 1) / 2);
    printf("oz_fwd %-6s P=%d sets=%d T=%d: %.4f ms per launch, %.0f int8 TOP/s; vs scaled: %zu of %zu values differ (max %.3g in units of the leading digit), %zu scales differ\n",
           names[mode], P, sets, T, ms / reps, ops / (ms / reps * 1e-3) / 1e12, bad, (size_t)rows * KpD, maxdiff, bad_s);
  }
  return 0;
}
