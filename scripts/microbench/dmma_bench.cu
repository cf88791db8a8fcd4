// Micro-benchmark: DMMA.8x8x4 (mma.sync m8n8k4 f64) throughput / latency against DFMA on one B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH> __global__ void k_dmma(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[CH][2];
  for(int i = 0; i < CH; ++i) c[i][0] = c[i][1] = i;
  for(int it = 0; it < iters; ++it) {
#pragma unroll
    for(int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for(int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH> __global__ void k_dfma(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[CH];
  for(int i = 0; i < CH; ++i) c[i] = i;
  for(int it = 0; it < iters; ++it) {
#pragma unroll
    for(int i = 0; i < CH; ++i) c[i] = fma(a, c[i], b);
  }
  double s = 0;
  for(int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double *out; cudaMalloc(&out, 148 * 1024 * 8 * sizeof(double));
  const int iters = 20000;
  int warps_list[] = {1, 2, 4, 8, 16, 32};
  for(int w : warps_list) {
    float ms = timeit([&] { k_dmma<8><<<148, 32 * w>>>(out, iters); });
    double fma = 148.0 * w * 8.0 * iters * 256.0;
    printf("DMMA CH=8 warps/SM=%2d: %.3f ms, %.2f TFLOP/s, %.2f clk per DMMA per SM (at 1.9 GHz)\n", w, ms, 2 * fma / ms * 1e-9, ms * 1e-3 * 1.9e9 / (w * 8.0 * iters));
  }
  for(int w : warps_list) {
    float ms = timeit([&] { k_dmma<1><<<148, 32 * w>>>(out, iters); });
    printf("DMMA CH=1 (dependent chain) warps/SM=%2d: %.3f ms, %.1f clk per DMMA per warp\n", w, ms, ms * 1e-3 * 1.9e9 / iters);
  }
  for(int w : warps_list) {
    float ms = timeit([&] { k_dfma<8><<<148, 32 * w>>>(out, iters); });
    double fma = 148.0 * w * 32.0 * 8.0 * iters;
    printf("DFMA CH=8 warps/SM=%2d: %.3f ms, %.2f TFLOP/s\n", w, ms, 2 * fma / ms * 1e-9);
  }
  return 0;
}
