// Developer micro-benchmark: issue rate of ONE warp on straight-line code vs a compact loop (instruction fetch).
#include <cstdio>
#include <cuda_runtime.h>
template <int N> struct Unroll { template <class F> static __device__ __forceinline__ void run(F f) { f(); Unroll<N - 1>::run(f); } };
template <> struct Unroll<0> { template <class F> static __device__ __forceinline__ void run(F) {} };
__global__ void k_straight(double* out, long long* cyc, int iters) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  const double b = 1.0000001, c = 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 64; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_loop(double* out, long long* cyc, int iters) {
  double a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  const double b = 1.0000001, c = 0.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int u = 0; u < 64; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_straight_f(float* out, long long* cyc, int iters) {
  float a[8]; for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  const float b = 1.0000001f, c = 0.5f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 64; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8); long long h; const int iters = 200;
  for (int thr : {32, 128, 512}) {
    k_straight<<<1, thr>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  straight-line 512 DFMA: %.2f cycles/instr\n", thr, (double)h / iters / 512);
    k_loop<<<1, thr>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  loop of 8 DFMA x64    : %.2f cycles/instr\n", thr, (double)h / iters / 512);
    k_straight_f<<<1, thr>>>((float*)out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %3d  straight-line 512 FFMA: %.2f cycles/instr\n", thr, (double)h / iters / 512);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
