// Microbenchmark: gemm_f64_kernel (emagls_b200/csrc/gemm.cuh) vs cuBLAS DGEMM on the shapes of the hot loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I emagls_b200/csrc -o /tmp/gemm_shapes \
//        tools/microbench/gemm_shapes.cu -lcublas && /tmp/gemm_shapes
#include <cstdio>
#include <functional>
#include <vector>
#include <cublas_v2.h>
#include "gemm.cuh"
using namespace emagls;
static float time_it(cudaStream_t st, int reps, const std::function<void()>& f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1, st); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}
int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  cublasHandle_t cb; cublasCreate(&cb); cublasSetStream(cb, st);
  struct Sh { int M, N, K; int akc, bkc; const char* name; };
  Sh shapes[] = {{2702, 14400, 400, 0, 1, "fwd                  "}, {14400, 400, 2702, 0, 1, "bwd                  "}};
  size_t maxel = 0;
  for (auto& s : shapes) { maxel = std::max(maxel, (size_t)s.M * s.K); maxel = std::max(maxel, (size_t)s.N * s.K); maxel = std::max(maxel, (size_t)s.M * s.N * 6); }
  double *A, *B, *C;
  cudaMalloc(&A, maxel * 8); cudaMalloc(&B, maxel * 8); cudaMalloc(&C, maxel * 8);
  cudaMemset(A, 0, maxel * 8); cudaMemset(B, 0, maxel * 8);
  for (auto& s : shapes) {
    GemmOperand a{A, s.akc ? (long long)s.K : (long long)s.M, s.akc}, b{B, s.bkc ? (long long)s.K : (long long)s.N, s.bkc};
    GemmShape sh{s.M, s.N, s.K};
    double fl = 2.0 * s.M * s.N * s.K;
    using Tall = GemmCfg<8, 5, 4, 2, 3, 1>;     // 256 x 80
    using Tall4 = GemmCfg<8, 5, 4, 2, 4, 1>;    // 256 x 80, 4 stages
    using Wide3 = GemmCfg<8, 4, 2, 4, 3, 1>;    // 128 x 128, 3 stages
    using Wide5 = GemmCfg<8, 4, 2, 4, 5, 1>;    // 128 x 128, 5 stages
    using W64 = GemmCfg<8, 2, 2, 4, 4, 1>;      // 128 x 64
    using T112 = GemmCfg<4, 7, 4, 2, 3, 1>;     // 128 x 112
    float t1 = time_it(st, 10, [&] { launch_gemm_cfg<GemmWide>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t3 = time_it(st, 10, [&] { launch_gemm_cfg<GemmNarrow>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t5 = time_it(st, 10, [&] { launch_gemm_cfg<Tall>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t6 = time_it(st, 10, [&] { launch_gemm_cfg<Tall4>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t7 = time_it(st, 10, [&] { launch_gemm_cfg<Wide3>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t8 = time_it(st, 10, [&] { launch_gemm_cfg<Wide5>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t9 = time_it(st, 10, [&] { launch_gemm_cfg<W64>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t10 = time_it(st, 10, [&] { launch_gemm_cfg<T112>(st, a, b, sh, EpiStore{C, s.N, 1.0}, 1); });
    float t11 = time_it(st, 10, [&] { launch_gemm_cfg<Tall>(st, a, b, sh, EpiStore{C, s.N, 1.0, (long long)s.M * s.N}, 2); });
    float t12 = time_it(st, 10, [&] { launch_gemm_cfg<Tall>(st, a, b, sh, EpiStore{C, s.N, 1.0, (long long)s.M * s.N}, 3); });
    printf("wide %.3f (%.1f) narrow %.3f (%.1f) tall256x80 %.3f (%.1f) tall4st %.3f (%.1f) wide3st %.3f (%.1f) wide5st %.3f (%.1f) 128x64 %.3f (%.1f) 128x112 %.3f (%.1f) tall-split2 %.3f (%.1f) tall-split3 %.3f (%.1f)\n",
      t1, fl/t1/1e9, t3, fl/t3/1e9, t5, fl/t5/1e9, t6, fl/t6/1e9, t7, fl/t7/1e9, t8, fl/t8/1e9, t9, fl/t9/1e9, t10, fl/t10/1e9, t11, fl/t11/1e9, t12, fl/t12/1e9);
    int ns = 1; float t2 = 0;
    // cuBLAS: C^T (N x M col-major) = op(B) op(A): use generic column-major call with the same flop count
    const double one = 1.0, zero = 0.0;
    float t4 = time_it(st, 10, [&] { cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, s.N, s.M, s.K, &one, B, s.K, A, s.K, &zero, C, s.N); });
    printf("%s M %5d N %5d K %5d | ours auto %7.3f ms %5.1f TF | split-K(%d) %7.3f ms %5.1f TF | narrow %7.3f ms %5.1f TF | cuBLAS %7.3f ms %5.1f TF\n",
           s.name, s.M, s.N, s.K, t1, fl / t1 / 1e9, ns, t2, fl / t2 / 1e9, t3, fl / t3 / 1e9, t4, fl / t4 / 1e9);
  }
  return 0;
}
