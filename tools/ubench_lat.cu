// Developer micro-benchmark: dependent-issue latencies (cycles) of the FP64 building blocks on B200.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#define LAT(name, init, body)                                                     \
  __global__ void k_##name(double* out, long long* cyc, int iters) {              \
    double x = 1.0000001 + threadIdx.x * 1e-9, y = 0.999999; init;                \
    long long t0 = clock64();                                                     \
    for (int i = 0; i < iters; ++i) { body; }                                     \
    long long t1 = clock64();                                                     \
    out[threadIdx.x] = x + y; if (threadIdx.x == 0) *cyc = t1 - t0;               \
  }
LAT(dfma, , x = fma(x, y, 1e-9))
LAT(dmul, , x = x * y)
LAT(dadd, , x = x + y)
LAT(rsqrt, , x = rsqrt(x) + 1.0)
LAT(drcp, , x = 1.0 / x + 0.5)
LAT(dsqrt, , x = sqrt(x) + 1.0)
LAT(shfl, , x = shfl_d(x, (threadIdx.x + 1) & 31))
LAT(dmma, double c1 = 0.0, asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(x), "+d"(c1) : "d"(y), "d"(y)))
LAT(ffma, float xf = 1.0f; float yf = 0.999f, xf = fmaf(xf, yf, 1e-9f); x = xf)
__global__ void k_lds(double* out, long long* cyc, int iters) {
  __shared__ int idx[64];
  idx[threadIdx.x] = (threadIdx.x + 1) & 31; __syncthreads();
  int p = threadIdx.x; long long t0 = clock64();
  for (int i = 0; i < iters; ++i) p = idx[p];
  long long t1 = clock64(); out[threadIdx.x] = p; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_sync(double* out, long long* cyc, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t1 = clock64(); if (threadIdx.x == 0) *cyc = t1 - t0; out[threadIdx.x] = 0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 8);
  const int iters = 4096; long long h;
#define RUN(name, thr) k_##name<<<1, thr>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-8s threads=%3d  %.1f cycles/op\n", #name, thr, (double)h / iters);
  RUN(dfma, 32) RUN(dmul, 32) RUN(dadd, 32) RUN(rsqrt, 32) RUN(drcp, 32) RUN(dsqrt, 32) RUN(shfl, 32) RUN(dmma, 32) RUN(ffma, 32) RUN(lds, 32)
  RUN(sync, 512) RUN(sync, 32) RUN(dfma, 512) RUN(dmma, 512)
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
