// Developer micro-benchmark: cycles of the 8x8 diagonal-block Cholesky + inverse (`chol8_inv`) used by
// ddqc_ddmpc.cu, alone and with other warps streaming DMMA updates (as in factor_columns).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/ubench_chol8 tools/ubench_chol8.cu
#include <cstdio>
#include "../data_driven_quad_control_b200/csrc/ddqc_ddmpc.cu"

namespace {
// mode bit 0..3: which warps stream DMMAs: 0 none, 1 all 15, 2 only the 12 on other sub-partitions, 3 only warps 4,8,12
// what: 0 empty loop, 1 chol8_inv, 2 load_diag only, 3 rsqrt chain of 8
__global__ void __launch_bounds__(512, 1) k_test(double* out, long long* cyc, int what, int mode, int iters) {
  extern __shared__ __align__(16) double sm[];
  __shared__ volatile int s_done;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int e = tid; e < 64 * 64; e += 512) sm[e] = ((e % 64) / 8 == (e % 64) % 8) ? 10.0 + (e & 7) : 0.01 * ((e * 7) % 13);
  if (tid == 0) s_done = 0;
  __syncthreads();
  if (tid < 64) { const int i = tid >> 3, c = tid & 7; sm[3072 + eoff(i, c)] = (i == c) ? 4.0 + i : 0.3 / (1 + (i > c ? i - c : c - i)); }
  __syncthreads();
  double acc_out = 0;
  if (warp == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      double* blk = sm + (it & 7) * 64;
      blk[lane] = sm[3072 + lane]; blk[32 + lane] = sm[3072 + 32 + lane];
      __syncwarp();
      if (what == 1) chol8_inv(blk, lane);
      if (what == 2) { double a[36]; load_diag(blk, a); acc_out += a[35] + a[17]; }
      if (what == 3) { double d = blk[lane]; for (int j = 0; j < 8; ++j) d = fma(d, rsqrt(d), 1.0); acc_out += d; }
      if (what == 4) { double d = blk[lane]; for (int j = 0; j < 72; ++j) d = fma(d, 0.999, 1.0); acc_out += d; }
      __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) { *cyc = t1 - t0; s_done = 1; }
  } else {
    const bool on = mode == 1 || (mode == 2 && (warp & 3) != 0) || (mode == 3 && (warp & 3) == 0);
    if (on) {
      double2 s0 = {0, 0}, s1 = {0, 0};
      while (!s_done) {
        double2 a, b;
        syrk_pair<false>(sm + 1024, 0, 4 + (warp & 3), 5 + (warp & 3), true, 3, 3, g, t, a, b);
        s0.x += a.x; s1.x += b.x;
      }
      acc_out = s0.x + s1.x;
    }
  }
  out[tid] = acc_out + sm[tid];
}
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 8);
  const int iters = 2000; long long h;
  const char* wn[] = {"empty", "chol8_inv", "load_diag", "8 dependent rsqrt+fma", "72 dependent fma"};
  const char* mn[] = {"alone", "15 warps DMMA", "12 warps DMMA (other sub-partitions)", "warps 4,8,12 DMMA (same sub-partition)"};
  for (int what = 0; what < 5; ++what)
    for (int mode = 0; mode < 4; ++mode) {
      k_test<<<1, 512, 64 * 64 * 8>>>(out, cyc, what, mode, iters);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-24s %-42s %7.0f cycles/call (%s)\n", wn[what], mn[mode], (double)h / iters, cudaGetErrorString(cudaGetLastError()));
    }
}
