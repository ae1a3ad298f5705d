// ddqc_ddmpc.cu -- batched nonlinear data-driven MPC solve for sm_100a.
//
// Replaces, for a batch of controllers, `NonlinearDataDrivenMPCController.
// update_and_solve_data_driven_mpc()` of the third-party direct_data_driven_mpc package the
// reference calls (data_driven_mpc/utilities/param_grid_search/controller_evaluation.py:361,
// ctor contract controller_creation.py:91-112): the Hankel-matrix QP over
// (alpha, sigma, ubar, ybar, u_s, y_s) of Berberich et al. 2022 (Part II), Problem (22).
//
// One persistent CTA per controller slot.  Instead of the 693-variable QP the kernel solves the
// algebraically equivalent condensed problem (derivation: DESIGN.md, numpy mirror:
// oracle/ddmpc_condensed.py):
//   1. G = H H^T, H = [H_u; 1^T; H_y] (r x C, r = (m+p)l + 1), built from the Hankel structure
//      (2l-1 sliding dot products per signal pair + a diagonal recurrence) -- H is never stored;
//   2. right-looking blocked Cholesky of the leading "hard" block (input rows + ones row), which
//      the two regularised systems share, leaving the output-row Schur complement S0;
//   3. alpha_sr enters only through v_sr = H alpha_sr = b_s - D_s (G + D_s)^-1 b_s
//      (tail Cholesky of S0 + D_s, one right-hand side);
//   4. tail Cholesky of S0 + D0, forward substitution X = L^-1 [K, c] for the nx+1 structured
//      right-hand sides, Z = X^T X  ->  box-constrained QP in nx = (L-n) m + m + p variables;
//   5. block principal pivoting on the box QP (exact; 1-5 small Cholesky solves in shared memory);
//   6. ubar_0..ubar_L, u_s, y_s out.
// Everything is fp64: cond(G) ~ 1e10-1e11 with lamb_sigma = 1e9 (measured), which rules out
// fp32 / TF32 / BF16 tensor-core arithmetic for the factorisation; see DESIGN.md for the
// tensor-core plan (FP64 DMMA for the trailing updates).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ddqc.h"

namespace {

constexpr int kThreads = 256;
constexpr int kNB = 32;        // Cholesky panel width
constexpr int kMaxR = 512;     // max rows of H
constexpr int kMaxNx = 128;    // max box-QP variables
constexpr int kMaxSig = 8;     // max m, p
constexpr int kMaxSlots = 296; // workspace slots (2 per SM)

struct Dims {
  int n, m, p, L, N, ell, C;
  int ru;   // hard block: m*ell input rows + ones row
  int ry;   // p*ell output rows
  int r;    // ru + ry
  int nB;   // (L-n)*m box rows
  int nx;   // nB + m + p
  int ld;   // leading dimension of the r x r workspace matrix
  int ldx;  // leading dimension (rows) of X
};

__host__ __device__ inline Dims make_dims(const DdqcMpcConfig& c) {
  Dims d;
  d.n = c.n; d.m = c.m; d.p = c.p; d.L = c.L; d.N = c.N;
  d.ell = c.L + c.n + 1;
  d.C = c.N - d.ell + 1;
  d.ru = c.m * d.ell + 1;
  d.ry = c.p * d.ell;
  d.r = d.ru + d.ry;
  d.nB = (c.L - c.n) * c.m;
  d.nx = d.nB + c.m + c.p;
  d.ld = (d.r + 3) & ~3;
  d.ldx = (d.r + 3) & ~3;
  return d;
}

__host__ __device__ inline size_t slot_doubles(const Dims& d) {
  // G/L (r x ld) | S copy (ry x ld) | X (ldx x (nx+1), column-major) | vectors (8 x ld)
  return (size_t)d.r * d.ld + (size_t)d.ry * d.ld + (size_t)d.ldx * (d.nx + 1) + (size_t)8 * d.ld;
}

// ---- small dense kernels on a lower-triangular matrix in global memory (row-major, ld) ------

// Unblocked Cholesky of the nb x nb diagonal block held in shared memory (row stride kNB+1),
// executed by warp 0.  Returns false on a non-positive pivot.
__device__ bool chol_diag_warp(double* sD, int nb, int lane) {
  bool ok = true;
  for (int j = 0; j < nb; ++j) {
    double djj = sD[j * (kNB + 1) + j];
    if (!(djj > 0.0)) ok = false;
    double dj = sqrt(djj);
    __syncwarp();
    if (lane == j) sD[j * (kNB + 1) + j] = dj;
    if (lane > j && lane < nb) sD[lane * (kNB + 1) + j] /= dj;
    __syncwarp();
    if (lane > j && lane < nb) {
      double lij = sD[lane * (kNB + 1) + j];
      for (int k = j + 1; k <= lane; ++k) sD[lane * (kNB + 1) + k] -= lij * sD[k * (kNB + 1) + j];
    }
    __syncwarp();
  }
  return ok;
}

// Right-looking blocked Cholesky of columns [c0, c1) of the n x n lower-triangular matrix A.
// After the call columns [c0, c1) hold L and the trailing block [c1, n) x [c1, n) holds the Schur
// complement.  sD: (kNB x (kNB+1)) doubles, sP: (n x kNB) doubles of shared memory.
__device__ bool chol_blocked(double* __restrict__ A, int ld, int n, int c0, int c1, double* sD, double* sP,
                             int* s_flag) {
  const int tid = threadIdx.x, lane = tid & 31;
  for (int k0 = c0; k0 < c1; k0 += kNB) {
    const int nb = min(kNB, c1 - k0);
    // diagonal block -> shared memory
    for (int e = tid; e < nb * nb; e += kThreads) {
      int i = e / nb, j = e % nb;
      sD[i * (kNB + 1) + j] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
    }
    __syncthreads();
    if (tid < 32) {
      bool ok = chol_diag_warp(sD, nb, lane);
      if (!ok && lane == 0) *s_flag = 1;
    }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += kThreads) {
      int i = e / nb, j = e % nb;
      if (j <= i) A[(size_t)(k0 + i) * ld + k0 + j] = sD[i * (kNB + 1) + j];
    }
    // panel: rows below the diagonal block, one thread per row (x kept in registers)
    const int rows = n - (k0 + nb);
    for (int rr = tid; rr < rows; rr += kThreads) {
      const int i = k0 + nb + rr;
      double x[kNB];
#pragma unroll
      for (int t = 0; t < kNB; ++t) x[t] = (t < nb) ? A[(size_t)i * ld + k0 + t] : 0.0;
#pragma unroll
      for (int t = 0; t < kNB; ++t) {
        if (t < nb) {
          double acc = x[t];
#pragma unroll
          for (int s = 0; s < t; ++s) acc -= x[s] * sD[t * (kNB + 1) + s];
          x[t] = acc / sD[t * (kNB + 1) + t];
        }
      }
#pragma unroll
      for (int t = 0; t < kNB; ++t) {
        if (t < nb) {
          A[(size_t)i * ld + k0 + t] = x[t];
          sP[rr * kNB + t] = x[t];
        }
      }
    }
    __syncthreads();
    // trailing update of the lower triangle: A[i][j] -= sum_t P[i][t] P[j][t], 4 x 4 tiles
    const int tiles = (rows + 3) / 4;
    const int ntile = tiles * (tiles + 1) / 2;
    for (int tt = tid; tt < ntile; tt += kThreads) {
      // unrank (ti >= tj) from tt
      int ti = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
      while ((ti + 1) * (ti + 2) / 2 <= tt) ++ti;
      while (ti * (ti + 1) / 2 > tt) --ti;
      const int tj = tt - ti * (ti + 1) / 2;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      const int i0 = ti * 4, j0 = tj * 4;
      for (int t = 0; t < nb; ++t) {
        double pi[4], pj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          pi[a] = (i0 + a < rows) ? sP[(i0 + a) * kNB + t] : 0.0;
          pj[a] = (j0 + a < rows) ? sP[(j0 + a) * kNB + t] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += pi[a] * pj[b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = i0 + a, j = j0 + b;
          if (i < rows && j <= i) A[(size_t)(k0 + nb + i) * ld + k0 + nb + j] -= acc[a][b];
        }
    }
    __syncthreads();
  }
  return true;
}

// Blocked forward substitution L X = B for `ncol` right-hand sides stored column-major in X (ldx
// rows), L lower-triangular n x n (row-major, ld).  Per row panel: a GEMM-shaped update with both
// operands contiguous (row of L, column of X), then the nb x nb triangular solve, one thread per
// column.
__device__ void forward_subst(const double* __restrict__ Lm, int ld, int n, double* __restrict__ X, int ldx, int ncol) {
  for (int k0 = 0; k0 < n; k0 += kNB) {
    const int nb = min(kNB, n - k0);
    if (k0 > 0) {
      for (int e = threadIdx.x; e < nb * ncol; e += kThreads) {
        const int i = k0 + e % nb, c = e / nb;
        const double* Li = Lm + (size_t)i * ld;
        const double* xc = X + (size_t)c * ldx;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = 0;
        for (; k + 3 < k0; k += 4) {
          a0 += Li[k] * xc[k];
          a1 += Li[k + 1] * xc[k + 1];
          a2 += Li[k + 2] * xc[k + 2];
          a3 += Li[k + 3] * xc[k + 3];
        }
        for (; k < k0; ++k) a0 += Li[k] * xc[k];
        X[(size_t)c * ldx + i] -= (a0 + a1) + (a2 + a3);
      }
      __syncthreads();
    }
    for (int c = threadIdx.x; c < ncol; c += kThreads) {
      double* xc = X + (size_t)c * ldx + k0;
      for (int t = 0; t < nb; ++t) {
        const double* Lt = Lm + (size_t)(k0 + t) * ld + k0;
        double acc = xc[t];
        for (int q = 0; q < t; ++q) acc -= Lt[q] * xc[q];
        xc[t] = acc / Lt[t];
      }
    }
    __syncthreads();
  }
}

// Solve (L L^T) x = b in place for ONE right-hand side with warp 0, where L is the composite
// factor [[L_uu, 0], [L_yu, L_t]]: rows < ru and the L_yu block live in A (ld), the tail factor L_t
// (ry x ry) in T (ldt).  Forward: row-oriented warp dot products; backward: column-oriented axpys
// over the (contiguous) rows of L.
__device__ void warp_solve_composite(const double* __restrict__ A, int ld, const double* __restrict__ T, int ldt, int ru,
                                     int r, double* __restrict__ x, int lane) {
  for (int i = 0; i < r; ++i) {
    const double* Li = (i < ru) ? A + (size_t)i * ld : A + (size_t)i * ld;
    const int nlead = i < ru ? i : ru;
    double acc = 0.0;
    for (int k = lane; k < nlead; k += 32) acc += Li[k] * x[k];
    if (i >= ru) {
      const double* Ti = T + (size_t)(i - ru) * ldt;
      for (int k = lane; k < i - ru; k += 32) acc += Ti[k] * x[ru + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const double diag = (i < ru) ? A[(size_t)i * ld + i] : T[(size_t)(i - ru) * ldt + (i - ru)];
    if (lane == 0) x[i] = (x[i] - acc) / diag;
    __syncwarp();
  }
  for (int k = r - 1; k >= 0; --k) {
    const double diag = (k < ru) ? A[(size_t)k * ld + k] : T[(size_t)(k - ru) * ldt + (k - ru)];
    const double xk = x[k] / diag;
    __syncwarp();
    if (lane == 0) x[k] = xk;
    // x_i -= L[k][i] * x_k for i < k  (row k of L)
    const double* Lk = A + (size_t)k * ld;
    const int nlead = k < ru ? k : ru;
    for (int i = lane; i < nlead; i += 32) x[i] -= Lk[i] * xk;
    if (k >= ru) {
      const double* Tk = T + (size_t)(k - ru) * ldt;
      for (int i = lane; i < k - ru; i += 32) x[ru + i] -= Tk[i] * xk;
    }
    __syncwarp();
  }
}

// ---- box-constrained strictly convex QP:  min 0.5 x'Hx + f'x,  lo <= x <= hi ------------------
//
// Stage 1 (whole CTA): W = H^-1 explicitly (Cholesky H = L L^T, L^-1, W = L^-T L^-1) and the
// unconstrained minimiser x0 = -W f.  Stage 2 (warp 0): block principal pivoting carried out in
// the space of the ACTIVE bounds only.  For an active set A with bound values b_A,
//     mu = W_AA^-1 (x0_A - b_A),   x = x0 - W[:,A] mu,     gradient_A = -mu,
// so a pass costs one |A| x |A| Cholesky (|A| is the handful of active bounds, not nx) plus an
// nx x |A| product.  All infeasible indices are exchanged at once; after three passes without
// progress the method switches to single exchanges of the largest infeasible index (Murty),
// which terminate finitely for an SPD Hessian.

// H (nx x nx, row-major, full) -> W = H^-1 in place; Li is scratch (nx x nx).
__device__ void spd_inverse_cta(double* H, double* Li, int nx, int* flag) {
  const int tid = threadIdx.x;
  for (int j = 0; j < nx; ++j) {
    if (tid == 0) {
      const double djj = H[j * nx + j];
      if (!(djj > 0.0)) *flag = 1;
      H[j * nx + j] = sqrt(djj);
    }
    __syncthreads();
    const double dj = H[j * nx + j];
    for (int i = j + 1 + tid; i < nx; i += kThreads) H[i * nx + j] /= dj;
    __syncthreads();
    const int rem = nx - j - 1;
    for (int e = tid; e < rem * rem; e += kThreads) {
      const int i = j + 1 + e / rem, k = j + 1 + e % rem;
      if (k <= i) H[i * nx + k] -= H[i * nx + j] * H[k * nx + j];
    }
    __syncthreads();
  }
  // Li = L^-1 (lower triangular), one thread per column: L y = e_c
  for (int c = tid; c < nx; c += kThreads) {
    for (int i = 0; i < c; ++i) Li[i * nx + c] = 0.0;
    for (int i = c; i < nx; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) acc -= H[i * nx + k] * Li[k * nx + c];
      Li[i * nx + c] = acc / H[i * nx + i];
    }
  }
  __syncthreads();
  // W = Li^T Li
  for (int e = tid; e < nx * nx; e += kThreads) {
    const int i = e / nx, k = e % nx;
    if (k > i) continue;
    double acc = 0.0;
    for (int t = i; t < nx; ++t) acc += Li[t * nx + i] * Li[t * nx + k];
    H[i * nx + k] = acc;
  }
  __syncthreads();
  for (int e = tid; e < nx * nx; e += kThreads) {
    const int i = e / nx, k = e % nx;
    if (k > i) H[i * nx + k] = H[k * nx + i];
  }
  __syncthreads();
}

// Stage 2, one warp.  W: nx x nx inverse Hessian; x0: unconstrained minimiser; F: scratch for the
// active block (nx x nx); state[i] in {0 free, -1 at lo, +1 at hi}; idx / verdict: int scratch.
__device__ void box_qp_bpp(const double* W, double* F, const double* x0, const double* lo, const double* hi, double* x,
                           double* mu, double tol_g, int* state, int* idx, int* verdict, int nx, int max_iter,
                           int lane, int* iters, int* converged, int* flag) {
  const double tol_x = 1e-10;
  int ninf_best = nx + 1, patience = 3;
  *converged = 0;
  int it = 0;
  for (; it < max_iter; ++it) {
    // ordered list of active indices
    int na = 0;
    for (int base = 0; base < nx; base += 32) {
      const int i = base + lane;
      const bool ac = i < nx && state[i] != 0;
      const unsigned mk = __ballot_sync(0xffffffffu, ac);
      if (ac) idx[na + __popc(mk & ((1u << lane) - 1u))] = i;
      na += __popc(mk);
    }
    __syncwarp();
    // F = W_AA (lower), mu = x0_A - b_A
    for (int e = lane; e < na * na; e += 32) {
      const int a = e / na, c = e % na;
      if (c <= a) F[a * nx + c] = W[idx[a] * nx + idx[c]];
    }
    for (int a = lane; a < na; a += 32) {
      const int i = idx[a];
      mu[a] = x0[i] - (state[i] < 0 ? lo[i] : hi[i]);
    }
    __syncwarp();
    for (int j = 0; j < na; ++j) {  // Cholesky of the active block
      const double djj = F[j * nx + j];
      if (!(djj > 0.0) && lane == 0) *flag = 1;
      const double dj = sqrt(djj);
      __syncwarp();
      if (lane == 0) F[j * nx + j] = dj;
      for (int i = j + 1 + lane; i < na; i += 32) F[i * nx + j] /= dj;
      __syncwarp();
      for (int i = j + 1 + lane; i < na; i += 32) {
        const double lij = F[i * nx + j];
        for (int k = j + 1; k <= i; ++k) F[i * nx + k] -= lij * F[k * nx + j];
      }
      __syncwarp();
    }
    for (int i = 0; i < na; ++i) {  // forward
      double acc = 0.0;
      for (int k = lane; k < i; k += 32) acc += F[i * nx + k] * mu[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) mu[i] = (mu[i] - acc) / F[i * nx + i];
      __syncwarp();
    }
    for (int i = na - 1; i >= 0; --i) {  // backward
      double acc = 0.0;
      for (int k = i + 1 + lane; k < na; k += 32) acc += F[k * nx + i] * mu[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) mu[i] = (mu[i] - acc) / F[i * nx + i];
      __syncwarp();
    }
    // x = x0 - W[:,A] mu ; active entries sit exactly on their bounds
    for (int i = lane; i < nx; i += 32) {
      double acc = x0[i];
      for (int a = 0; a < na; ++a) acc -= W[i * nx + idx[a]] * mu[a];
      x[i] = state[i] < 0 ? lo[i] : (state[i] > 0 ? hi[i] : acc);
    }
    __syncwarp();
    // verdicts.  gradient_i = -mu_a: at a lower bound it must be >= 0, at an upper bound <= 0.
    for (int i = lane; i < nx; i += 32) verdict[i] = 0;
    __syncwarp();
    for (int a = lane; a < na; a += 32) {
      const int i = idx[a];
      const double gi = -mu[a];
      if ((state[i] < 0 && gi < -tol_g) || (state[i] > 0 && gi > tol_g)) verdict[i] = 2;
    }
    for (int i = lane; i < nx; i += 32) {
      if (state[i] == 0) {
        if (x[i] < lo[i] - tol_x * (1.0 + fabs(lo[i]))) verdict[i] = -1;
        else if (x[i] > hi[i] + tol_x * (1.0 + fabs(hi[i]))) verdict[i] = 1;
      }
    }
    __syncwarp();
    int ninf = 0, last = -1;
    for (int base = 0; base < nx; base += 32) {
      const int i = base + lane;
      const unsigned mk = __ballot_sync(0xffffffffu, i < nx && verdict[i] != 0);
      ninf += __popc(mk);
      if (mk) last = base + 31 - __clz(mk);
    }
    if (ninf == 0) {
      *converged = 1;
      ++it;
      break;
    }
    bool full;
    if (ninf < ninf_best) {
      ninf_best = ninf;
      patience = 3;
      full = true;
    } else if (patience > 0) {
      --patience;
      full = true;
    } else {
      full = false;
    }
    for (int i = lane; i < nx; i += 32) {
      if (!full && i != last) continue;
      const int v = verdict[i];
      if (v == -1) state[i] = -1;
      else if (v == 1) state[i] = 1;
      else if (v == 2) state[i] = 0;
    }
    __syncwarp();
  }
  *iters = it;
}

// ---- the solver ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
ddmpc_solve_kernel(const DdqcMpcConfig cfg, double* __restrict__ ws, const double* __restrict__ u_win,
                   const double* __restrict__ y_win, const double* __restrict__ y_r,
                   const double* __restrict__ us_prev, const double* __restrict__ ys_prev,
                   const int32_t* __restrict__ have_prev, double* __restrict__ u_opt, double* __restrict__ us_opt,
                   double* __restrict__ ys_opt, int32_t* __restrict__ status, int32_t* __restrict__ iters,
                   const int batch, long long* __restrict__ phase_clk) {
  extern __shared__ double smem[];
  const Dims d = make_dims(cfg);
  const int tid = threadIdx.x;
  const int m = d.m, p = d.p, ell = d.ell, C = d.C, N = d.N, n = d.n, L = d.L;
  const int nsig = m + p;

  // shared memory carve-up
  double* sW = smem;                              // windows, signal-major: nsig x N
  double* sD = sW + (size_t)nsig * N;             // kNB x (kNB+1)
  double* sv = sD + kNB * (kNB + 1);              // vectors: f, lo, hi, x, g, rhs (6 x kMaxNx)
  double* sP = sv + 6 * kMaxNx;                   // Cholesky panel: r x kNB ...
  double* sH = sP;                                // ... aliased later by the box-QP Hessian nx x nx
  double* sF = sH + (size_t)d.nx * d.nx;          // and the factor of its free block
  __shared__ int s_flag;
  __shared__ int s_state[kMaxNx];
  __shared__ int s_idx[kMaxNx];
  __shared__ int s_verdict[kMaxNx];
  __shared__ int s_cnt[4];

  double* slot = ws + (size_t)blockIdx.x * slot_doubles(d);
  double* Gm = slot;                              // r x ld
  double* S2 = Gm + (size_t)d.r * d.ld;           // ry x ld  (Schur complement copy for alpha_sr)
  double* X = S2 + (size_t)d.ry * d.ld;           // ldx x (nx+1), column-major
  double* vec = X + (size_t)d.ldx * (d.nx + 1);   // 8 x ld scratch vectors
  double* v_sr = vec;                             // r
  double* cs = vec + d.ld;                        // r

  // row index helpers (row order: input rows j*m+i | ones row | output rows)
  const int row_one = m * ell;
  const int row_y0 = m * ell + 1;

  long long t_prev = clock64();
#define DDQC_PHASE(k)                                                  \
  do {                                                                 \
    if (phase_clk && tid == 0) {                                       \
      const long long t_now = clock64();                               \
      atomicAdd((unsigned long long*)&phase_clk[k], (unsigned long long)(t_now - t_prev)); \
      t_prev = t_now;                                                  \
    }                                                                  \
  } while (0)
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    if (tid == 0) s_flag = 0;
    t_prev = clock64();
    // ---- stage the (u, y) window, signal-major ----
    const double* ub = u_win + (size_t)b * N * m;
    const double* yb = y_win + (size_t)b * N * p;
    for (int e = tid; e < N * m; e += kThreads) sW[(e % m) * N + e / m] = ub[e];
    for (int e = tid; e < N * p; e += kThreads) sW[(m + e % p) * N + e / p] = yb[e];
    __syncthreads();

    // ---- 1. Hankel Gram, lower triangle of G ----
    // G[(j,a),(k,b)] = sum_c w_a[j+c] w_b[k+c]; one task per (a, b, diagonal delta = j-k in (-ell, ell))
    {
      const int ndiag = 2 * ell - 1;
      const int ntask = nsig * nsig * ndiag;
      for (int t = tid; t < ntask; t += kThreads) {
        const int a = t / (nsig * ndiag), bsig = (t / ndiag) % nsig, dl = t % ndiag - (ell - 1);
        int j = dl > 0 ? dl : 0, k = dl > 0 ? 0 : -dl;
        const double* wa = sW + a * N;
        const double* wb = sW + bsig * N;
        double f = 0.0;
        for (int c = 0; c < C; ++c) f += wa[j + c] * wb[k + c];
        for (; j < ell && k < ell; ++j, ++k) {
          const int ra = (a < m) ? j * m + a : row_y0 + j * p + (a - m);
          const int rb = (bsig < m) ? k * m + bsig : row_y0 + k * p + (bsig - m);
          if (rb <= ra) Gm[(size_t)ra * d.ld + rb] = f;
          if (j + 1 >= ell || k + 1 >= ell) break;
          f += wa[j + C] * wb[k + C] - wa[j] * wb[k];  // j + C <= N - 1 here
        }
      }
      // ones row: sliding sums, and G[one][one] = C
      for (int t = tid; t < nsig; t += kThreads) {
        const double* wa = sW + t * N;
        double f = 0.0;
        for (int c = 0; c < C; ++c) f += wa[c];
        for (int j = 0; j < ell; ++j) {
          const int ra = (t < m) ? j * m + t : row_y0 + j * p + (t - m);
          if (ra > row_one) Gm[(size_t)ra * d.ld + row_one] = f;
          else Gm[(size_t)row_one * d.ld + ra] = f;
          if (j + 1 < ell) f += wa[j + C] - wa[j];
        }
      }
      if (tid == 0) Gm[(size_t)row_one * d.ld + row_one] = (double)C;
    }
    __syncthreads();

    DDQC_PHASE(0);
    // ---- 2. shared factorisation of the hard block (columns [0, ru)) ----
    chol_blocked(Gm, d.ld, d.r, 0, d.ru, sD, sP, &s_flag);
    // trailing block [ru, r) now holds S0; copy it for the alpha_sr system
    const bool use_sr = (cfg.alpha_reg_type == 0) && have_prev && have_prev[b] != 0;
    if (use_sr) {
      for (int e = tid; e < d.ry * d.ry; e += kThreads) {
        int i = e / d.ry, j = e % d.ry;
        if (j <= i) S2[(size_t)i * d.ld + j] = Gm[(size_t)(d.ru + i) * d.ld + d.ru + j];
      }
    }
    __syncthreads();

    DDQC_PHASE(1);
    // ---- 3. v_sr = b_s - D_s (G + D_s)^-1 b_s ----
    for (int i = tid; i < d.r; i += kThreads) v_sr[i] = 0.0;
    if (use_sr) {
      const double ds = cfg.lamb_alpha_s / cfg.lamb_sigma_s;
      for (int i = tid; i < d.ry; i += kThreads) S2[(size_t)i * d.ld + i] += ds;
      __syncthreads();
      // S2 is a ry x ry matrix of its own (offset 0)
      chol_blocked(S2, d.ld, d.ry, 0, d.ry, sD, sP, &s_flag);
      // b_s
      for (int i = tid; i < d.r; i += kThreads) {
        double v;
        if (i < row_one) v = us_prev[(size_t)b * m + i % m];
        else if (i == row_one) v = 1.0;
        else v = ys_prev[(size_t)b * p + (i - row_y0) % p];
        cs[i] = v;
      }
      __syncthreads();
      if (tid < 32) warp_solve_composite(Gm, d.ld, S2, d.ld, d.ru, d.r, cs, tid);
      __syncthreads();
      for (int i = tid; i < d.r; i += kThreads) {
        double bs;
        if (i < row_one) bs = us_prev[(size_t)b * m + i % m];
        else if (i == row_one) bs = 1.0;
        else bs = ys_prev[(size_t)b * p + (i - row_y0) % p];
        v_sr[i] = (i >= row_y0) ? bs - ds * cs[i] : bs;
      }
    }
    __syncthreads();

    DDQC_PHASE(2);
    // ---- 4. tail factorisation with D0, right-hand sides, Z = X^T X ----
    for (int i = tid; i < d.ry; i += kThreads) {
      const int j = i / p, comp = i % p;
      double w;
      if (j < n || j >= L) w = cfg.lamb_sigma;
      else w = cfg.Q[comp] * cfg.lamb_sigma / (cfg.Q[comp] + cfg.lamb_sigma);
      Gm[(size_t)(d.ru + i) * d.ld + d.ru + i] += cfg.lamb_alpha / w;
    }
    __syncthreads();
    chol_blocked(Gm, d.ld, d.r, d.ru, d.r, sD, sP, &s_flag);

    DDQC_PHASE(3);
    // K = [E_B | T] and c = tau - v_sr, column-major into X
    for (int e = tid; e < d.ldx * (d.nx + 1); e += kThreads) X[e] = 0.0;
    __syncthreads();
    for (int c = tid; c < d.nx + 1; c += kThreads) {
      double* x = X + (size_t)c * d.ldx;
      if (c < d.nB) {
        x[n * m + c] = 1.0;  // box row of predicted input (j = n + c / m, component c % m)
      } else if (c < d.nB + m) {
        const int comp = c - d.nB;
        for (int j = L; j < ell; ++j) x[j * m + comp] = 1.0;  // terminal input rows = u_s
      } else if (c < d.nx) {
        const int comp = c - d.nB - m;
        for (int j = n; j < ell; ++j) x[row_y0 + j * p + comp] = 1.0;  // free + terminal output rows target y_s
      } else {
        for (int i = 0; i < d.r; ++i) {
          double tau = 0.0;
          if (i < n * m) tau = ub[(size_t)(N - n) * m + i];  // past inputs
          else if (i == row_one) tau = 1.0;
          else if (i >= row_y0 && i < row_y0 + n * p) tau = yb[(size_t)(N - n) * p + (i - row_y0)];
          x[i] = tau - v_sr[i];
        }
      }
    }
    __syncthreads();
    forward_subst(Gm, d.ld, d.r, X, d.ldx, d.nx + 1);
    __syncthreads();

    DDQC_PHASE(4);
    // Z = X^T X (nx+1 columns): Hq = 2 lamb_alpha Z[:nx,:nx] + structure, f = 2 lamb_alpha Z[:nx, nx] + ...
    double* sf = sv;
    double* slo = sv + kMaxNx;
    double* shi = sv + 2 * kMaxNx;
    double* sx = sv + 3 * kMaxNx;
    double* sg = sv + 4 * kMaxNx;
    double* srhs = sv + 5 * kMaxNx;
    {
      const int nc = d.nx + 1;
      for (int e = tid; e < d.nx * nc; e += kThreads) {
        const int a = e / nc, c = e % nc;
        if (c < d.nx && c > a) continue;  // lower triangle + the rhs column
        const double* xa = X + (size_t)a * d.ldx;
        const double* xc = X + (size_t)c * d.ldx;
        double acc = 0.0;
        for (int i = 0; i < d.r; ++i) acc += xa[i] * xc[i];
        acc *= 2.0 * cfg.lamb_alpha;
        if (c < d.nx) {
          sH[a * d.nx + c] = acc;
          sH[c * d.nx + a] = acc;
        } else {
          sf[a] = acc;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      // structured part of the box-QP objective
      for (int bi = 0; bi < d.nB; ++bi) {
        const int comp = bi % m;
        const double w2 = 2.0 * cfg.R[comp];
        sH[bi * d.nx + bi] += w2;
        sH[bi * d.nx + d.nB + comp] -= w2;
        sH[(d.nB + comp) * d.nx + bi] -= w2;
        sH[(d.nB + comp) * d.nx + d.nB + comp] += w2;
      }
      for (int i = 0; i < m; ++i) {
        double su = 0.0;
        for (int j = 0; j < n; ++j) su += ub[(size_t)(N - n + j) * m + i];
        sH[(d.nB + i) * d.nx + d.nB + i] += 2.0 * n * cfg.R[i];
        sf[d.nB + i] += -2.0 * cfg.R[i] * su;
      }
      for (int i = 0; i < p; ++i) {
        double sy = 0.0;
        for (int j = 0; j < n; ++j) sy += yb[(size_t)(N - n + j) * p + i];
        const int q = d.nB + m + i;
        sH[q * d.nx + q] += 2.0 * (n * cfg.Q[i] + cfg.S[i]);
        sf[q] += -2.0 * (cfg.Q[i] * sy + cfg.S[i] * y_r[(size_t)b * p + i]);
      }
    }
    for (int i = tid; i < d.nx; i += kThreads) {
      double lo, hi;
      if (i < d.nB) {
        lo = cfg.U_lo[i % m];
        hi = cfg.U_hi[i % m];
      } else if (i < d.nB + m) {
        const int comp = i - d.nB;
        lo = fmax(cfg.U_lo[comp], cfg.Us_lo[comp]);
        hi = fmin(cfg.U_hi[comp], cfg.Us_hi[comp]);
      } else {
        lo = -INFINITY;
        hi = INFINITY;
      }
      slo[i] = lo;
      shi[i] = hi;
      s_state[i] = 0;
    }
    __syncthreads();

    DDQC_PHASE(5);
    // ---- 5. box QP: explicit inverse Hessian (CTA) + block principal pivoting on the active set (warp 0) ----
    {
      double fmax_abs = 0.0;
      for (int i = 0; i < d.nx; ++i) fmax_abs = fmax(fmax_abs, fabs(sf[i]));
      const double tol_g = 1e-11 * fmax(1.0, fmax_abs);
      spd_inverse_cta(sH, sF, d.nx, &s_flag);
      for (int i = tid; i < d.nx; i += kThreads) {  // x0 = -W f
        double acc = 0.0;
        for (int k = 0; k < d.nx; ++k) acc -= sH[i * d.nx + k] * sf[k];
        sg[i] = acc;
      }
      __syncthreads();
      DDQC_PHASE(6);
      if (tid < 32) {
        int it_w = 0, conv_w = 0;
        box_qp_bpp(sH, sF, sg, slo, shi, sx, srhs, tol_g, s_state, s_idx, s_verdict, d.nx,
                   cfg.max_iter > 0 ? cfg.max_iter : 20000, tid, &it_w, &conv_w, &s_flag);
        if (tid == 0) {
          s_cnt[2] = it_w;
          s_cnt[3] = conv_w;
        }
      }
      __syncthreads();
    }
    const int it = s_cnt[2], converged = s_cnt[3];

    DDQC_PHASE(7);
    // ---- 6. outputs ----
    for (int e = tid; e < (L + 1) * m; e += kThreads) {
      const int k = e / m, comp = e % m;
      u_opt[(size_t)b * (L + 1) * m + e] = (k < L - n) ? sx[k * m + comp] : sx[d.nB + comp];
    }
    for (int i = tid; i < m; i += kThreads) us_opt[(size_t)b * m + i] = sx[d.nB + i];
    for (int i = tid; i < p; i += kThreads) ys_opt[(size_t)b * p + i] = sx[d.nB + m + i];
    if (tid == 0) {
      status[b] = s_flag ? 2 : (converged ? 0 : 1);
      iters[b] = it;
    }
    __syncthreads();
  }
}

__global__ void ddmpc_store_kernel(const int N, const int m, const int p, double* __restrict__ u_win,
                                   double* __restrict__ y_win, const double* __restrict__ u_new,
                                   const double* __restrict__ y_new, const int batch) {
  // one CTA per controller: shift the window one sample (in place, ordered) and append
  const int b = blockIdx.x;
  if (b >= batch) return;
  double* ub = u_win + (size_t)b * N * m;
  double* yb = y_win + (size_t)b * N * p;
  // in-place shift needs read-before-write ordering: go through registers in chunks of the CTA
  for (int base = 0; base < (N - 1) * m; base += blockDim.x) {
    const int e = base + threadIdx.x;
    double v = (e < (N - 1) * m) ? ub[e + m] : 0.0;
    __syncthreads();
    if (e < (N - 1) * m) ub[e] = v;
    __syncthreads();
  }
  for (int base = 0; base < (N - 1) * p; base += blockDim.x) {
    const int e = base + threadIdx.x;
    double v = (e < (N - 1) * p) ? yb[e + p] : 0.0;
    __syncthreads();
    if (e < (N - 1) * p) yb[e] = v;
    __syncthreads();
  }
  if (threadIdx.x < m) ub[(size_t)(N - 1) * m + threadIdx.x] = u_new[(size_t)b * m + threadIdx.x];
  if (threadIdx.x < p) yb[(size_t)(N - 1) * p + threadIdx.x] = y_new[(size_t)b * p + threadIdx.x];
}

int check_cfg(const DdqcMpcConfig* c) {
  if (!c) return DDQC_E_NULL;
  if (c->m < 1 || c->p < 1 || c->m > kMaxSig || c->p > kMaxSig || c->n < 1 || c->L <= c->n) return DDQC_E_SHAPE;
  const Dims d = make_dims(*c);
  if (d.C < 1 || d.r > kMaxR || d.nx > kMaxNx) return DDQC_E_SHAPE;
  if (c->alpha_reg_type != 0 && c->alpha_reg_type != 2) return DDQC_E_CONFIG;
  return 0;
}

size_t smem_bytes(const Dims& d) {
  const size_t panel = (size_t)d.r * kNB, boxqp = 2 * (size_t)d.nx * d.nx;
  size_t dbl = (size_t)(d.m + d.p) * d.N + kNB * (kNB + 1) + 6 * kMaxNx + (panel > boxqp ? panel : boxqp);
  return dbl * sizeof(double);
}

}  // namespace

static long long* g_phase_clk = nullptr;
// Developer aid (not part of the reference-facing surface): device buffer of 8 cycle counters that the
// solver accumulates per phase (Gram, hard-block Cholesky, alpha_sr, tail Cholesky, substitutions, Z,
// inverse, pivoting).  Pass NULL to switch it off.
extern "C" int ddqc_ddmpc_debug_phase_buffer(void* dev_ptr) {
  g_phase_clk = static_cast<long long*>(dev_ptr);
  return 0;
}

extern "C" int64_t ddqc_ddmpc_workspace_bytes(const DdqcMpcConfig* cfg, int64_t batch) {
  if (check_cfg(cfg) != 0 || batch < 0) return -1;
  const Dims d = make_dims(*cfg);
  const int64_t slots = batch < kMaxSlots ? batch : kMaxSlots;
  return (int64_t)(slot_doubles(d) * sizeof(double)) * (slots > 0 ? slots : 1);
}

extern "C" int ddqc_ddmpc_solve(const DdqcMpcConfig* cfg, void* ws, const double* u_win, const double* y_win,
                                const double* y_r, const double* us_prev, const double* ys_prev,
                                const int32_t* have_prev, double* u_opt, double* us_opt, double* ys_opt,
                                int32_t* status, int32_t* iters, int64_t batch, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!ws || !u_win || !y_win || !y_r || !u_opt || !us_opt || !ys_opt || !status || !iters) return DDQC_E_NULL;
  if (cfg->alpha_reg_type == 0 && (!us_prev || !ys_prev || !have_prev)) return DDQC_E_NULL;
  if (batch < 0) return DDQC_E_SHAPE;
  if (batch == 0) return 0;
  const Dims d = make_dims(*cfg);
  const size_t smem = smem_bytes(d);
  if (smem > 227 * 1024) return DDQC_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(ddmpc_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = (int)(batch < sms ? batch : sms);
  if (grid > kMaxSlots) grid = kMaxSlots;
  ddmpc_solve_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      *cfg, static_cast<double*>(ws), u_win, y_win, y_r, us_prev, ys_prev, have_prev, u_opt, us_opt, ys_opt, status,
      iters, (int)batch, g_phase_clk);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_ddmpc_store(const DdqcMpcConfig* cfg, double* u_win, double* y_win, const double* u_new,
                                const double* y_new, int64_t batch, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!u_win || !y_win || !u_new || !y_new) return DDQC_E_NULL;
  if (batch <= 0) return batch == 0 ? 0 : DDQC_E_SHAPE;
  ddmpc_store_kernel<<<(unsigned)batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(cfg->N, cfg->m, cfg->p, u_win, y_win,
                                                                                    u_new, y_new, (int)batch);
  return (int)cudaGetLastError();
}
