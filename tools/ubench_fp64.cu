// Developer micro-benchmark: FP64 throughput of one B200 SM pipe, DFMA vs DMMA (mma.sync.m8n8k4.f64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
  double a[8], b = 1.0000001, c = 0.5;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(double) * 4);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {128, 256, 512, 1024}) {
    for (int which = 0; which < 2; ++which) {
      int iters = 20000;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_dfma<<<sms, threads>>>(out, iters); else k_dmma<<<sms, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma_per_thread = which == 0 ? 8.0 * iters : 8.0 * iters * 8.0;  // m8n8k4 = 256 FMA / 32 lanes
      double tflops = 2.0 * fma_per_thread * threads * sms / (ms * 1e-3) / 1e12;
      printf("%s threads/SM=%4d  %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", which ? "DMMA" : "DFMA", threads, ms, tflops,
             tflops * 1e12 / 2 / sms / 1.965e9);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
