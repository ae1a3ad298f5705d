// Developer micro-benchmark: cycles of the 8x8 diagonal-block Cholesky / triangular solve used by
// ddqc_ddmpc.cu, alone and with the other 15 warps streaming DMMA updates (as in factor_columns).
#include <cstdio>
#include "../data_driven_quad_control_b200/csrc/ddqc_ddmpc.cu"

namespace {
// V3: every lane factors the whole block redundantly in registers
__device__ __forceinline__ bool chol8_redundant(double* blk, int lane) {
  double a[36];
  load_diag(blk, a);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double dj = LT(j, j);
    if (!(dj > 0.0)) { ok = false; dj = 1.0; }
    const double r = rsqrt(dj);
    LT(j, j) = r;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) LT(i, j) *= r;
#pragma unroll
    for (int c = j + 1; c < 8; ++c)
#pragma unroll
      for (int i = c; i < 8; ++i) LT(i, c) -= LT(i, j) * LT(c, j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c <= i; ++c)
      if (lane == ((i * (i + 1) / 2 + c) & 31)) blk[eoff(i, c)] = LT(i, c);
  return ok;
}
// V4: like V3 but with a hand-rolled reciprocal square root (fp32 seed + 2 Newton steps + 1 correction)
__device__ __forceinline__ double fast_rsqrt(double d) {
  double r = (double)rsqrtf((float)d);
  r = r * (1.5 - 0.5 * d * r * r);
  r = r * (1.5 - 0.5 * d * r * r);
  return r;
}
__device__ __forceinline__ bool chol8_redundant_fast(double* blk, int lane) {
  double a[36];
  load_diag(blk, a);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double dj = LT(j, j);
    if (!(dj > 0.0)) { ok = false; dj = 1.0; }
    const double r = fast_rsqrt(dj);
    LT(j, j) = r;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) LT(i, j) *= r;
#pragma unroll
    for (int c = j + 1; c < 8; ++c)
#pragma unroll
      for (int i = c; i < 8; ++i) LT(i, c) -= LT(i, j) * LT(c, j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c <= i; ++c)
      if (lane == ((i * (i + 1) / 2 + c) & 31)) blk[eoff(i, c)] = LT(i, c);
  return ok;
}
}

namespace {
template <int VAR>
__device__ __forceinline__ bool chol8_var(double* blk, int lane) {
  const int r0 = lane >> 3, c = lane & 7;
  double A = blk[eoff(r0, c)], B = blk[eoff(r0 + 4, c)];
  bool ok = true;
  double dj = shfl_d(A, 0);
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    if (!(dj > 0.0)) { ok = false; dj = 1.0; }
    const double r = (VAR & 1) ? dj * 0.3 : rsqrt(dj);
    if (c == j) {
      A = (r0 == j) ? r : A * r;
      B = (r0 + 4 == j) ? r : B * r;
    }
    const int jn = (j + 1) & 7, ln = (jn & 3) << 3;
    double x1, x2, y1, y2, ln_j, dn;
    if (VAR & 2) { x1 = A; x2 = B; y1 = A * 0.5; y2 = B * 0.5; ln_j = A * 0.25; dn = B + 1.0; }
    else {
      x1 = shfl_d(A, (r0 << 3) | j); x2 = shfl_d(B, (r0 << 3) | j);
      y1 = shfl_d(A, ((c & 3) << 3) | j); y2 = shfl_d(B, ((c & 3) << 3) | j);
      ln_j = shfl_d(jn < 4 ? A : B, ln | j);
      dn = shfl_d(jn < 4 ? A : B, ln | jn);
    }
    const double lc = c < 4 ? y1 : y2;
    if (c > j) {
      if (r0 >= c) A -= x1 * lc;
      if (r0 + 4 >= c) B -= x2 * lc;
    }
    dj = dn - ln_j * ln_j;
  }
  if (r0 >= c) blk[eoff(r0, c)] = A;
  if (r0 + 4 >= c) blk[eoff(r0 + 4, c)] = B;
  return ok;
}
}
namespace {
__global__ void __launch_bounds__(512, 1) k_test(double* out, long long* cyc, int mode, int iters) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int e = tid; e < 64 * 64; e += 512) sm[e] = ((e % 64) / 8 == (e % 64) % 8) ? 10.0 + (e & 7) : 0.01 * ((e * 7) % 13);
  __syncthreads();
  // pristine SPD block (diagonally dominant) kept at sm + 3072, rhs at sm + 3136
  if (tid < 64) { const int i = tid >> 3, c = tid & 7; sm[3072 + eoff(i, c)] = (i == c) ? 4.0 + i : 0.3 / (1 + (i > c ? i - c : c - i)); sm[3136 + tid] = 0.1 * tid; }
  __syncthreads();
  long long t0 = clock64();
  double acc_out = 0;
  if (warp == 0) {
    for (int it = 0; it < iters; ++it) {
      sm[(it & 7) * 64 + lane] = sm[3072 + lane]; sm[(it & 7) * 64 + 32 + lane] = sm[3072 + 32 + lane];
      sm[512 + (it & 7) * 64 + lane] = sm[3136 + lane]; sm[512 + (it & 7) * 64 + 32 + lane] = sm[3136 + 32 + lane];
      __syncwarp();
      if (mode == 1 || mode == 5) { chol8_warp(sm + (it & 7) * 64, lane); }
      if (mode == 32) chol8_var<0>(sm + (it & 7) * 64, lane);
      if (mode == 33) chol8_var<1>(sm + (it & 7) * 64, lane);
      if (mode == 34) chol8_var<2>(sm + (it & 7) * 64, lane);
      if (mode == 35) chol8_var<3>(sm + (it & 7) * 64, lane);
      if (mode == 8) { chol8_redundant(sm + (it & 7) * 64, lane); }
      if (mode == 9 || mode == 13) { double aa[36]; chol8_regs(sm + (it & 7) * 64, aa); acc_out += aa[35]; }
      if (mode == 16) { chol8_redundant_fast(sm + (it & 7) * 64, lane); }
      if (mode == 2 || mode == 6) { double a[36]; load_diag(sm + (it & 7) * 64, a); double2 v = ld_c(sm + 512 + (it & 7) * 64, g, t); trsm8_frag(v, a, t); st_c(sm + 512 + (it & 7) * 64, g, t, v); }
      // restore a positive diagonal so that repeated factorisation stays finite
      if (lane < 8) sm[(it & 7) * 64 + eoff(lane, lane)] = 10.0 + lane;
      __syncwarp();
    }
  } else if (mode == 5 || mode == 6 || (mode == 13 && (warp & 3) != 0)) {
    double2 s0 = {0, 0}, s1 = {0, 0};
    for (int it = 0; it < iters * 4; ++it) {
      double2 a, b;
      syrk_pair<false>(sm + 1024, 0, 4 + (warp & 3), 5 + (warp & 3), true, 3, 3, g, t, a, b);
      s0.x += a.x; s1.x += b.x;
    }
    acc_out = s0.x + s1.x;
  }
  long long t1 = clock64();
  if (tid == 0) *cyc = t1 - t0;
  out[tid] = acc_out + sm[tid];
}
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 * 8);
  const int iters = 2000; long long h;
  for (int mode : {99, 1, 9, 13, 2}) {
    k_test<<<1, 512, 64 * 64 * 8>>>(out, cyc, mode, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mode %2d (99 empty; 1 chol8 shuffle; 9 chol8 registers; 13 registers + DMMA on the other 3 sub-partitions; 2 trsm) %.0f cycles/call   (%s)\n", mode, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
  }
}
