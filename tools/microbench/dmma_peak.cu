// Microbenchmark: sustained DMMA.8x8x4 rate on sm_100a as a function of warps per SM and of the
// number of independent accumulators per warp (no memory traffic).  Build & run under gpurun:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_peak tools/microbench/dmma_peak.cu && /tmp/dmma_peak
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dmma_loop(double* out, int iters, double seed) {
  double acc[NACC][2];
  for (int i = 0; i < NACC; ++i) { acc[i][0] = seed * i; acc[i][1] = seed; }
  double a = seed + threadIdx.x, b = seed - threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(int warps_per_sm, double* out) {
  int iters = 20000 / NACC * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dmma_loop<NACC><<<148, warps_per_sm * 32>>>(out, 10, 1.0);
  cudaEventRecord(e0);
  dmma_loop<NACC><<<148, warps_per_sm * 32>>>(out, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 148.0 * warps_per_sm * (double)iters * NACC * 512.0;
  printf("warps/SM %2d  acc/warp %2d : %7.2f TFLOP/s\n", warps_per_sm, NACC, flops / (ms * 1e-3) / 1e12);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
  for (int w : {4, 8, 16, 32}) { run<4>(w, out); run<8>(w, out); run<16>(w, out); run<32>(w, out); }
  return 0;
}
