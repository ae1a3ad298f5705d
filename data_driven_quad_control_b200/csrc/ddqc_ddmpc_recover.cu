// ddqc_ddmpc_recover.cu -- the two optional modes of the third-party NonlinearDataDrivenMPCController that need
// the optimal alpha, which the condensed solver (ddqc_ddmpc.cu) never forms:
//
//   alpha_reg_type = 1 (PREVIOUS)  "regularized w.r.t. a previous optimal alpha value"
//                                  configs/data_driven_mpc/dd_mpc_controller_params.yaml:60-66
//   update_cost_threshold          "online input-output data updates are disabled when the tracking cost value is
//                                  less than this value"                         dd_mpc_controller_params.yaml:85-90
//   (ctor keywords: data_driven_mpc/utilities/param_grid_search/controller_creation.py:108-110)
//
// Dual recovery.  With the box-QP optimum (ubar, u_s, y_s) of a solve known, every row of H = [H_u; H_y; 1'] has a
// prescribed value t_i (input rows, ones row: hard, D_i = 0) or a target t_i with weight w_i (output rows: soft,
// D_i = lamb_alpha / w_i), and the minimising alpha is alpha_sr + H'z with
//
//        (G + D) z = t - v_sr,       G = H H',  v_sr = H alpha_sr.
//
// From z:  alpha* = alpha_sr + H'z;  ybar_k - y_s = -(lamb_alpha / Q) z on the predicted output rows, hence the
// tracking cost  sum_k |ubar_k - u_s|_R^2 + |ybar_k - y_s|_Q^2 + |y_s - y_r|_S^2.   numpy mirror:
// oracle/ddmpc_condensed.py::recover_dual (checked against the literal-QP oracle on the CPU).
//
// One 256-thread CTA per controller (persistent, work counter).  G is generated from the Hankel structure (one dot
// product per signal pair and diagonal, then the two-term recurrence along the diagonal) into a per-CTA workspace
// matrix that stays in L2; blocked right-looking Cholesky (16 columns per step: diagonal block by one warp in shared
// memory, panel rows one per thread, trailing update in 4x4 register tiles against the panel held column-major in
// shared memory); the right-hand side rides along as an extra row, so the forward substitution is free; blocked
// back-substitution.  These modes are off in every shipped configuration: the kernel is written for clarity, and
// costs about 1.7 times the solve it follows (ncu: 39 % of the warp samples wait at the barriers around the one-warp
// diagonal-block factorisation).  Measured and dropped: a LEFT-looking factorisation that keeps each row's block-column
// accumulator in registers and never writes the trailing matrix back (every thread streams its own row of L from the
// workspace: 14.2 ms per 4096 PREVIOUS solves against 13.3 ms - the strided row reads cost more than the write-backs).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ddqc.h"

namespace {

constexpr int kT = 256;
constexpr int kNB = 16;
#ifndef DDQC_REC_BLOCKS
#define DDQC_REC_BLOCKS 3  // resident CTAs per SM the kernel is compiled for: 2 -> 14.8 ms per 4096 PREVIOUS solves (solve
#endif                     // + recovery), 3 -> 13.3, 4 -> 13.4 (profiles/r2_ddmpc_recover_variants.log)
constexpr int kMaxSlotsR = 640;
constexpr int kSmemLimitR = 227 * 1024;

struct RDims {
  int n, m, p, L, N, ell, C, nsig, r, rp, ld;
};

__host__ __device__ inline RDims make_rdims(const DdqcMpcConfig& c) {
  RDims d;
  d.n = c.n; d.m = c.m; d.p = c.p; d.L = c.L; d.N = c.N;
  d.ell = c.L + c.n + 1;
  d.C = c.N - d.ell + 1;
  d.nsig = c.m + c.p;
  d.r = d.nsig * d.ell + 1;                 // Hankel rows: signal a, depth j at a * ell + j; ones row last
  d.rp = (d.r + kNB - 1) / kNB * kNB;       // padded with identity rows to whole blocks
  d.ld = d.rp + kNB;                        // row rp is the right-hand side
  return d;
}

__host__ __device__ inline size_t rslot_doubles(const RDims& d) { return (size_t)(d.rp + 1) * d.ld; }
__host__ __device__ inline size_t rsmem_doubles(const RDims& d) {
  return (((size_t)d.nsig * d.N + 1) & ~(size_t)1) + (size_t)kNB * d.ld + kNB * (kNB + 1) + d.ld + 64;
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kT / 32; ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(kT, DDQC_REC_BLOCKS)
ddmpc_recover_kernel(const DdqcMpcConfig cfg, double* __restrict__ ws, int* __restrict__ work_counter,
                     const double* __restrict__ u_win, const double* __restrict__ y_win,
                     const double* __restrict__ u_ic, const double* __restrict__ y_ic, const double* __restrict__ y_r,
                     const int32_t* __restrict__ have_prev, const double* __restrict__ lamb,
                     const double* __restrict__ u_opt, const double* __restrict__ us_opt,
                     const double* __restrict__ ys_opt, const double* __restrict__ v_sr, const double* __restrict__ sr,
                     double* __restrict__ alpha_io, double* __restrict__ cost_out, const int batch) {
  extern __shared__ __align__(16) double smr[];
  const RDims d = make_rdims(cfg);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = d.n, m = d.m, p = d.p, L = d.L, N = d.N, ell = d.ell, C = d.C, nsig = d.nsig, r = d.r, rp = d.rp, ld = d.ld;
  double* sW = smr;                        // window, signal-major [nsig][N]
  double* sP = sW + (((size_t)nsig * N + 1) & ~(size_t)1);  // panel, column-major [kNB][ld] (row index relative to the panel's first row)
  double* sD = sP + (size_t)kNB * ld;      // diagonal block [kNB][kNB + 1]; its diagonal holds 1 / l_cc
  double* sz = sD + kNB * (kNB + 1);       // y, then z
  double* s_red = sz + ld;
  __shared__ int s_b, s_bad;
  double* A = ws + (size_t)blockIdx.x * rslot_doubles(d);

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_b = atomicAdd(work_counter, 1);
      s_bad = 0;
    }
    __syncthreads();
    const int b = s_b;
    if (b >= batch) break;
    if (have_prev[b] < 0) continue;
    const double lam_a = lamb ? lamb[4 * (size_t)b] : cfg.lamb_alpha, lam_s = lamb ? lamb[4 * (size_t)b + 1] : cfg.lamb_sigma;
    const double* ub = u_win + (size_t)b * N * m;
    const double* yb = y_win + (size_t)b * N * p;
    const double* ubp = u_ic ? u_ic + (size_t)b * N * m : ub;
    const double* ybp = y_ic ? y_ic + (size_t)b * N * p : yb;
    const double* uo = u_opt + (size_t)b * (L + 1) * m;
    const double* us = us_opt + (size_t)b * m;
    const double* ys = ys_opt + (size_t)b * p;

    for (int e = tid; e < N * m; e += kT) sW[(e % m) * N + e / m] = ub[e];
    for (int e = tid; e < N * p; e += kT) sW[(m + e % p) * N + e / p] = yb[e];
    __syncthreads();

    // ---- G (lower triangle; whole blocks for the signal pairs a >= bs) ----------------------------------------------
    {
      const int ndiag = 2 * ell - 1, npair = nsig * (nsig + 1) / 2;
      for (int task = tid; task < npair * ndiag; task += kT) {
        const int pr = task / ndiag, dg = task % ndiag;
        int a = (int)((sqrtf(8.0f * pr + 1.0f) - 1.0f) * 0.5f);
        while ((a + 1) * (a + 2) / 2 <= pr) ++a;
        while (a * (a + 1) / 2 > pr) --a;
        const int bs = pr - a * (a + 1) / 2;
        const int i0 = dg < ell ? 0 : dg - ell + 1, j0 = dg < ell ? dg : 0;
        const double* wa = sW + a * N + i0;
        const double* wb = sW + bs * N + j0;
        double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
        int c = 0;
        for (; c + 3 < C; c += 4) {
          g0 = fma(wa[c], wb[c], g0);
          g1 = fma(wa[c + 1], wb[c + 1], g1);
          g2 = fma(wa[c + 2], wb[c + 2], g2);
          g3 = fma(wa[c + 3], wb[c + 3], g3);
        }
        for (; c < C; ++c) g0 = fma(wa[c], wb[c], g0);
        double gv = (g0 + g1) + (g2 + g3);
        const int len = ell - (i0 > j0 ? i0 : j0);
        double* dst = A + (size_t)(a * ell + i0) * ld + (bs * ell + j0);
        for (int k = 0; k < len; ++k) {
          dst[(size_t)k * (ld + 1)] = gv;
          if (k + 1 < len) gv = fma(wa[k + C], wb[k + C], fma(-wa[k], wb[k], gv));
        }
      }
      // ones row
      for (int q = tid; q < nsig * ell; q += kT) {
        const double* w = sW + (q / ell) * N + q % ell;
        double s0 = 0.0, s1 = 0.0;
        int c = 0;
        for (; c + 1 < C; c += 2) { s0 += w[c]; s1 += w[c + 1]; }
        if (c < C) s0 += w[c];
        A[(size_t)(r - 1) * ld + q] = s0 + s1;
      }
      if (tid == 0) A[(size_t)(r - 1) * ld + (r - 1)] = (double)C;
      // identity pad rows
      for (int e = tid; e < (rp - r) * rp; e += kT) {
        const int i = r + e / rp, j = e % rp;
        if (j <= i) A[(size_t)i * ld + j] = i == j ? 1.0 : 0.0;
      }
    }
    __syncthreads();
    // ---- D on the diagonal, right-hand side t - v_sr as row rp --------------------------------------------------------
    for (int q = tid; q < rp; q += kT) {
      double rhs = 0.0;
      if (q < r - 1) {
        const int a = q / ell, j = q % ell;
        double t, vs = 0.0;
        if (a < m) {
          t = j < n ? ubp[(size_t)(N - n + j) * m + a] : uo[(size_t)(j - n) * m + a];
          if (sr) vs = sr[(size_t)b * (nsig + p * ell) + a];
        } else {
          const int c = a - m;
          t = j < n ? ybp[(size_t)(N - n + j) * p + c] : ys[c];
          const double w = (j < n || j >= L) ? lam_s : cfg.Q[c] * lam_s / (cfg.Q[c] + lam_s);
          A[(size_t)q * ld + q] += lam_a / w;
          if (sr) vs = sr[(size_t)b * (nsig + p * ell) + a] - sr[(size_t)b * (nsig + p * ell) + nsig + j * p + c];
        }
        if (v_sr) vs = v_sr[(size_t)b * r + q];
        rhs = t - vs;
      } else if (q == r - 1) {
        rhs = 1.0 - (v_sr ? v_sr[(size_t)b * r + q] : (sr ? 1.0 : 0.0));
      }
      A[(size_t)rp * ld + q] = rhs;
    }
    __syncthreads();

    // ---- blocked Cholesky, rows 0..rp (row rp = right-hand side) ------------------------------------------------------
    for (int k0 = 0; k0 < rp; k0 += kNB) {
      const int k1 = k0 + kNB;
      {
        const int i = tid / kNB, j = tid % kNB;  // kT == kNB * kNB
        if (j <= i) sD[i * (kNB + 1) + j] = A[(size_t)(k0 + i) * ld + k0 + j];
      }
      __syncthreads();
      if (warp == 0) {
        for (int c = 0; c < kNB; ++c) {
          const double dcc = sD[c * (kNB + 1) + c];
          if (!(dcc > 0.0) && lane == 0) s_bad = 1;
          const double inv = rsqrt(dcc);
          __syncwarp();
          if (lane == c) sD[c * (kNB + 1) + c] = inv;
          if (lane > c && lane < kNB) sD[lane * (kNB + 1) + c] *= inv;
          __syncwarp();
          if (lane > c && lane < kNB) {
            const double lic = sD[lane * (kNB + 1) + c];
            for (int c2 = c + 1; c2 <= lane; ++c2) sD[lane * (kNB + 1) + c2] -= lic * sD[c2 * (kNB + 1) + c];
          }
          __syncwarp();
        }
      }
      __syncthreads();
      {  // factor of the diagonal block back to the matrix (diagonal entries as 1 / l_cc) for the back-substitution
        const int i = tid / kNB, j = tid % kNB;
        if (j <= i) A[(size_t)(k0 + i) * ld + k0 + j] = sD[i * (kNB + 1) + j];
      }
      // panel: one row per thread, X <- X L_kk^-T; kept column-major in shared memory for the trailing update
      for (int i = k1 + tid; i <= rp; i += kT) {
        double x[kNB];
        double* row = A + (size_t)i * ld + k0;
#pragma unroll
        for (int c = 0; c < kNB; ++c) x[c] = row[c];
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
          double s = x[c];
#pragma unroll
          for (int c2 = 0; c2 < c; ++c2) s = fma(-x[c2], sD[c * (kNB + 1) + c2], s);
          x[c] = s * sD[c * (kNB + 1) + c];
        }
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
          row[c] = x[c];
          sP[(size_t)c * ld + (i - k1)] = x[c];
        }
      }
      __syncthreads();
      // trailing update: rows [k1, rp], columns [k1, min(row, rp - 1)], 4x4 tiles
      {
        const int nrow = rp + 1 - k1, nt = (nrow + 3) >> 2, ntask = nt * (nt + 1) / 2;
        for (int task = tid; task < ntask; task += kT) {
          int ti = (int)((sqrtf(8.0f * task + 1.0f) - 1.0f) * 0.5f);
          while ((ti + 1) * (ti + 2) / 2 <= task) ++ti;
          while (ti * (ti + 1) / 2 > task) --ti;
          const int tj = task - ti * (ti + 1) / 2;
          // the tile's current values are requested FIRST, as double2 rows, so that their L2 latency runs under the
          // panel product (read-modify-write statements on the same array are not reordered by the compiler: as sixteen
          // `A[..] -= acc` the tile cost sixteen dependent round trips)
          double2 cur[4][2];
          double* tp = A + (size_t)(k1 + 4 * ti) * ld + (k1 + 4 * tj);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const bool in = k1 + 4 * ti + x <= rp;
            cur[x][0] = in ? *reinterpret_cast<const double2*>(tp + (size_t)x * ld) : make_double2(0.0, 0.0);
            cur[x][1] = in ? *reinterpret_cast<const double2*>(tp + (size_t)x * ld + 2) : make_double2(0.0, 0.0);
          }
          double acc[4][4];
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
          const double* pi = sP + 4 * ti;
          const double* pj = sP + 4 * tj;
#pragma unroll 4
          for (int c = 0; c < kNB; ++c) {
            const double2 a01 = *reinterpret_cast<const double2*>(pi + (size_t)c * ld);
            const double2 a23 = *reinterpret_cast<const double2*>(pi + (size_t)c * ld + 2);
            const double2 b01 = *reinterpret_cast<const double2*>(pj + (size_t)c * ld);
            const double2 b23 = *reinterpret_cast<const double2*>(pj + (size_t)c * ld + 2);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
          }
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int i = k1 + 4 * ti + x;
            if (i > rp) continue;
            const double cv[4] = {cur[x][0].x, cur[x][0].y, cur[x][1].x, cur[x][1].y};
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              const int j = k1 + 4 * tj + y;
              if (j <= i && j < rp) tp[(size_t)x * ld + y] = cv[y] - acc[x][y];
            }
          }
        }
      }
      __syncthreads();
    }

    // ---- z = L^-T y, y = row rp ------------------------------------------------------------------------------------------
    for (int q = tid; q < rp; q += kT) sz[q] = A[(size_t)rp * ld + q];
    __syncthreads();
    for (int k0 = rp - kNB; k0 >= 0; k0 -= kNB) {
      {
        const int i = tid / kNB, j = tid % kNB;
        if (j <= i) sD[i * (kNB + 1) + j] = A[(size_t)(k0 + i) * ld + k0 + j];  // diagonal entries hold 1 / l_cc
      }
      __syncthreads();
      if (tid == 0) {
        for (int c = kNB - 1; c >= 0; --c) {
          double s = sz[k0 + c];
          for (int c2 = c + 1; c2 < kNB; ++c2) s = fma(-sD[c2 * (kNB + 1) + c], sz[k0 + c2], s);
          sz[k0 + c] = s * sD[c * (kNB + 1) + c];
        }
      }
      __syncthreads();
      for (int j = tid; j < k0; j += kT) {
        double s = sz[j];
#pragma unroll 4
        for (int l = 0; l < kNB; ++l) s = fma(-A[(size_t)(k0 + l) * ld + j], sz[k0 + l], s);
        sz[j] = s;
      }
      __syncthreads();
    }

    // ---- outputs ------------------------------------------------------------------------------------------------------
    const bool bad = s_bad != 0;
    if (cost_out) {
      double acc = 0.0;
      for (int e = tid; e < ell * m; e += kT) {
        const int j = e / m, a = e % m;
        const double v = (j < n ? ubp[(size_t)(N - n + j) * m + a] : uo[(size_t)(j - n) * m + a]) - us[a];
        acc = fma(cfg.R[a] * v, v, acc);
      }
      for (int e = tid; e < n * p; e += kT) {
        const int j = e / p, c = e % p;
        const double v = ybp[(size_t)(N - n + j) * p + c] - ys[c];
        acc = fma(cfg.Q[c] * v, v, acc);
      }
      for (int e = tid; e < (L - n) * p; e += kT) {
        const int j = n + e / p, c = e % p;
        const double v = lam_a * sz[(m + c) * ell + j];  // ybar - y_s = -(lamb_alpha / Q) z
        acc = fma(v / cfg.Q[c], v, acc);
      }
      if (tid < p) {
        const double v = ys[tid] - y_r[(size_t)b * p + tid];
        acc = fma(cfg.S[tid] * v, v, acc);
      }
      const double tot = block_sum(acc, s_red);
      if (tid == 0) cost_out[b] = bad ? INFINITY : tot;
    }
    if (alpha_io && !bad) {
      double* al = alpha_io + (size_t)b * C;
      for (int c = tid; c < C; c += kT) {
        double s0 = sz[r - 1], s1 = 0.0;
        for (int a = 0; a < nsig; ++a) {
          const double* w = sW + a * N + c;
          const double* za = sz + a * ell;
          int j = 0;
          for (; j + 1 < ell; j += 2) {
            s0 = fma(w[j], za[j], s0);
            s1 = fma(w[j + 1], za[j + 1], s1);
          }
          if (j < ell) s0 = fma(w[j], za[j], s0);
        }
        al[c] += s0 + s1;
      }
    }
  }
}

// v = H alpha: one CTA per controller
__global__ void __launch_bounds__(kT)
ddmpc_hankel_apply_kernel(const DdqcMpcConfig cfg, const double* __restrict__ u_win, const double* __restrict__ y_win,
                          const double* __restrict__ alpha, double* __restrict__ v, const int batch) {
  extern __shared__ __align__(16) double smr[];
  const RDims d = make_rdims(cfg);
  const int b = blockIdx.x, tid = threadIdx.x;
  if (b >= batch) return;
  const int m = d.m, p = d.p, N = d.N, ell = d.ell, C = d.C, nsig = d.nsig, r = d.r;
  double* sW = smr;
  double* sa = sW + (size_t)nsig * N;
  const double* ub = u_win + (size_t)b * N * m;
  const double* yb = y_win + (size_t)b * N * p;
  for (int e = tid; e < N * m; e += kT) sW[(e % m) * N + e / m] = ub[e];
  for (int e = tid; e < N * p; e += kT) sW[(m + e % p) * N + e / p] = yb[e];
  for (int c = tid; c < C; c += kT) sa[c] = alpha[(size_t)b * C + c];
  __syncthreads();
  for (int q = tid; q < r; q += kT) {
    double s0 = 0.0, s1 = 0.0;
    if (q < r - 1) {
      const double* w = sW + (q / ell) * N + q % ell;
      int c = 0;
      for (; c + 1 < C; c += 2) {
        s0 = fma(w[c], sa[c], s0);
        s1 = fma(w[c + 1], sa[c + 1], s1);
      }
      if (c < C) s0 = fma(w[c], sa[c], s0);
    } else {
      for (int c = 0; c < C; ++c) s0 += sa[c];
    }
    v[(size_t)b * r + q] = s0 + s1;
  }
}

__global__ void __launch_bounds__(kT)
ddmpc_refresh_data_kernel(const DdqcMpcConfig cfg, const double* __restrict__ u_roll, const double* __restrict__ y_roll,
                          double* __restrict__ u_data, double* __restrict__ y_data, const double* __restrict__ cost,
                          const double threshold, const int32_t* __restrict__ have_prev, int32_t* __restrict__ refreshed,
                          const int batch) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  const int hp = have_prev[b];
  if (hp < 0) return;
  const bool keep = hp > 0 && cost[b] <= threshold;  // NaN / inf cost: refresh
  if (threadIdx.x == 0 && refreshed) refreshed[b] = keep ? 0 : 1;
  if (keep) return;
  const size_t nu = (size_t)cfg.N * cfg.m, ny = (size_t)cfg.N * cfg.p;
  for (size_t e = threadIdx.x; e < nu; e += kT) u_data[b * nu + e] = u_roll[b * nu + e];
  for (size_t e = threadIdx.x; e < ny; e += kT) y_data[b * ny + e] = y_roll[b * ny + e];
}

int check_rcfg(const DdqcMpcConfig* c) {
  if (!c) return DDQC_E_NULL;
  if (c->m < 1 || c->p < 1 || c->m > 8 || c->p > 8 || c->n < 1 || c->L <= c->n) return DDQC_E_SHAPE;
  const RDims d = make_rdims(*c);
  if (d.C < 1) return DDQC_E_SHAPE;
  if (rsmem_doubles(d) * sizeof(double) + 1024 > (size_t)kSmemLimitR) return DDQC_E_SHAPE;
  return 0;
}

}  // namespace

extern "C" int64_t ddqc_ddmpc_recover_workspace_bytes(const DdqcMpcConfig* cfg, int64_t batch) {
  if (check_rcfg(cfg) != 0 || batch < 0) return -1;
  const RDims d = make_rdims(*cfg);
  const int64_t slots = batch < kMaxSlotsR ? (batch > 0 ? batch : 1) : kMaxSlotsR;
  return 256 + (int64_t)(rslot_doubles(d) * sizeof(double)) * slots;
}

extern "C" int ddqc_ddmpc_recover(const DdqcMpcConfig* cfg, void* ws, const double* u_win, const double* y_win,
                                  const double* u_ic, const double* y_ic, const double* y_r, const int32_t* have_prev,
                                  const double* lamb, const double* u_opt, const double* us_opt, const double* ys_opt,
                                  const double* v_sr, const double* sr, double* alpha_io, double* cost_out,
                                  int64_t batch, void* stream) {
  int rc = check_rcfg(cfg);
  if (rc) return rc;
  if (!ws || !u_win || !y_win || !y_r || !have_prev || !u_opt || !us_opt || !ys_opt) return DDQC_E_NULL;
  if ((u_ic == nullptr) != (y_ic == nullptr)) return DDQC_E_NULL;
  if (!alpha_io && !cost_out) return DDQC_E_NULL;
  // what alpha_sr was must match the mode the solve ran in
  if (cfg->alpha_reg_type == 1 && (!v_sr || sr)) return DDQC_E_NULL;
  if (cfg->alpha_reg_type == 0 && (!sr || v_sr)) return DDQC_E_NULL;
  if (cfg->alpha_reg_type == 2 && (sr || v_sr)) return DDQC_E_CONFIG;
  if (batch < 0) return DDQC_E_SHAPE;
  if (batch == 0) return 0;
  const RDims d = make_rdims(*cfg);
  const size_t smem = rsmem_doubles(d) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(ddmpc_recover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = DDQC_REC_BLOCKS * sms;
  if (grid > kMaxSlotsR) grid = kMaxSlotsR;
  if (grid > batch) grid = (int)batch;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  e = cudaMemsetAsync(ws, 0, 256, st);
  if (e != cudaSuccess) return (int)e;
  double* slots = reinterpret_cast<double*>(static_cast<char*>(ws) + 256);
  ddmpc_recover_kernel<<<grid, kT, smem, st>>>(*cfg, slots, static_cast<int*>(ws), u_win, y_win, u_ic, y_ic, y_r,
                                               have_prev, lamb, u_opt, us_opt, ys_opt, v_sr, sr, alpha_io, cost_out,
                                               (int)batch);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_ddmpc_hankel_apply(const DdqcMpcConfig* cfg, const double* u_win, const double* y_win,
                                       const double* alpha, double* v, int64_t batch, void* stream) {
  int rc = check_rcfg(cfg);
  if (rc) return rc;
  if (!u_win || !y_win || !alpha || !v) return DDQC_E_NULL;
  if (batch <= 0) return batch == 0 ? 0 : DDQC_E_SHAPE;
  const RDims d = make_rdims(*cfg);
  const size_t smem = ((size_t)d.nsig * d.N + d.C) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(ddmpc_hankel_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ddmpc_hankel_apply_kernel<<<(unsigned)batch, kT, smem, static_cast<cudaStream_t>(stream)>>>(*cfg, u_win, y_win, alpha, v,
                                                                                              (int)batch);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_ddmpc_refresh_data(const DdqcMpcConfig* cfg, const double* u_roll, const double* y_roll,
                                       double* u_data, double* y_data, const double* cost, double threshold,
                                       const int32_t* have_prev, int32_t* refreshed, int64_t batch, void* stream) {
  int rc = check_rcfg(cfg);
  if (rc) return rc;
  if (!u_roll || !y_roll || !u_data || !y_data || !cost || !have_prev) return DDQC_E_NULL;
  if (batch <= 0) return batch == 0 ? 0 : DDQC_E_SHAPE;
  ddmpc_refresh_data_kernel<<<(unsigned)batch, kT, 0, static_cast<cudaStream_t>(stream)>>>(
      *cfg, u_roll, y_roll, u_data, y_data, cost, threshold, have_prev, refreshed, (int)batch);
  return (int)cudaGetLastError();
}
