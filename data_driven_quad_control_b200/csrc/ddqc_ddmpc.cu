// ddqc_ddmpc.cu -- batched nonlinear data-driven MPC solve for sm_100a.
//
// Replaces, for a batch of controllers, `NonlinearDataDrivenMPCController.
// update_and_solve_data_driven_mpc()` of the third-party direct_data_driven_mpc package the
// reference calls (data_driven_mpc/utilities/param_grid_search/controller_evaluation.py:361,
// ctor contract controller_creation.py:91-112): the Hankel-matrix QP over
// (alpha, sigma, ubar, ybar, u_s, y_s) of Berberich et al. 2022 (Part II), Problem (22).
//
// One persistent 512-thread CTA per SM pulls controllers from a work counter.  The 693-variable
// QP is condensed (derivation: DESIGN.md section 4; numpy mirror: oracle/ddmpc_condensed.py) into
// Schur complements of G + D, G = H H^T the Gram matrix of H = [H_u; 1^T; H_y], and solved as a
// box-constrained QP in nx = (L-n) m + m + p variables.  Per controller:
//
//   0. Gram.  G is generated from the Hankel structure (one sliding dot product per signal pair
//      and diagonal, then a two-term recurrence) and scattered, already permuted into the
//      elimination order [R1 | R2 | R3 | aug], into an L2-resident workspace slot.
//        R1 = past-input rows, ones row, terminal-input rows (hard, shared by both systems)
//        R2 = output rows (soft: diagonal regularisation D differs between the two systems)
//        R3 = predicted-input rows (the box-constrained decision variables y_B)
//        aug = right-hand sides carried as extra rows: A_u (m), A_y (p), e_one, tau | T_u (m), T_y (p)
//      Closed loop (optional per-controller Gram cache in HBM): when the window moved by one sample since the cached
//      G, step 0 is the rank-2 correction G' = G - a a' + b b' (a = old first, b = new last Hankel column) instead,
//      fused with step A: cached panel blocks are corrected where the panel is factored, trailing blocks stream
//      through two shared-memory stages straight into the S1 update (see "Gram cache" below).
//   A. Eliminate R1 (left-looking blocked Cholesky of the panel in shared memory), then
//      S1 = trailing - L L^T streamed through global memory once.  S1 is used twice:
//   B. alpha_sr system (alpha of the optimal reachable equilibrium for y_r): S1 + D_s with the aug rows
//      A_u, A_y, e_one is factored completely; the aug x aug Schur block gives a 6-variable box QP for
//      (u_sr, y_sr); rho = A_u u_sr + A_y y_sr + e_one is forward-substituted for free (same combination
//      of the aug rows), a blocked back-substitution gives e = D_s (G + D_s)^-1 rho on R2 (v_sr = rho - e).
//   C. main system: S1 + D_0, aug rows T_u, T_y and c = tau - rho + e, eliminate R2 only; the Schur
//      complement on [R3 | aug] is eliminated once more with an identity block appended (tensor cores), which yields
//      Z = K'(G + D_0)^-1 K, i.e. the box-QP Hessian and gradient.
//   D. box QP: block principal pivoting in the space of the FREE variables, warm-started from
//      the previous solve's active set (shifted one step); primal-dual interior point fallback
//      followed by a pivoting polish when pivoting does not settle.
//
// The whole trailing matrix lives in shared memory (203 KB at the default sizes) in 8x8 blocks
// (block-packed lower triangle, XOR-4 column swizzle inside a block) so that every FP64 tensor
// core fragment (`mma.sync.m8n8k4.f64`, DMMA) is loaded without bank conflicts.  All Gram /
// Schur contractions run on DMMA; fp64 is required (cond(G) ~ 1e10, D_0 entries of 5e-8).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ddqc.h"

namespace {

#ifndef DDQC_MPC_THREADS
#define DDQC_MPC_THREADS 512
#endif
constexpr int kThreads = DDQC_MPC_THREADS;
#ifndef DDQC_BULK_ALL
#define DDQC_BULK_ALL 1  // 1: warps 4, 8, 12 (warp 0's SM sub-partition) also take part in the lookahead updates
#endif
constexpr int kWarps = kThreads / 32;
constexpr int kMaxNx = 128;    // max box-QP variables
constexpr int kMaxSig = 8;     // max m, p
constexpr int kMaxSlots = 160; // workspace slots (>= SM count)
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxPos = 1024;  // max (m + p) * ell
constexpr int kRepackQ = (((kMaxNx / 8 + 2) * (kMaxNx / 8 + 3) / 2) * 64 + kThreads - 1) / kThreads;  // Schur repack, values per thread
constexpr int kAsmRows = (kMaxNx + 1 + kWarps - 1) / kWarps, kAsmCols = (kMaxNx + 31) / 32;               // Hq/f assembly: rows per warp, columns per lane

struct Dims {
  int n, m, p, L, N, ell, C;
  int n1, n2, n3, na;        // class sizes (rows)
  int n1p, n2p, n3p;         // padded to multiples of 8
  int nb1, nb2, nb3;         // in 8x8 blocks
  int nbt, nbA;              // trailing block rows, all block rows
  int nx, nS, ldS, ldq;
  int ga, gs, gsg;           // doubles: R1 panel (rect), trailing in shared memory / in the workspace (block-packed lower)
};

__host__ __device__ inline int rup8(int v) { return (v + 7) & ~7; }

__host__ __device__ inline Dims make_dims(const DdqcMpcConfig& c) {
  Dims d;
  d.n = c.n; d.m = c.m; d.p = c.p; d.L = c.L; d.N = c.N;
  d.ell = c.L + c.n + 1;
  d.C = c.N - d.ell + 1;
  d.n1 = c.m * (2 * c.n + 1) + 1;
  d.n2 = c.p * d.ell;
  d.n3 = (c.L - c.n) * c.m;
  d.na = c.m + c.p + 2;  // aug group 0: A_u, A_y, e_one, tau (group 1: T_u, T_y); each fits one block row
  d.n1p = rup8(d.n1); d.n2p = rup8(d.n2); d.n3p = rup8(d.n3);
  d.nb1 = d.n1p / 8; d.nb2 = d.n2p / 8; d.nb3 = d.n3p / 8;
  d.nbt = d.nb2 + d.nb3 + 1;      // block rows held in shared memory (matrix + ONE aug block row)
  d.nbA = d.nb1 + d.nbt + 1;      // block rows of the R1 panel (both aug block rows)
  d.nx = d.n3 + c.m + c.p;
  d.nS = d.nx + 1;
  d.ldS = (d.nS + 1) & ~1;  // even: rows are accessed as double2
  d.ldq = d.nx | 1;
  d.ga = d.nbA * d.nb1 * 64;
  d.gs = d.nbt * (d.nbt + 1) / 2 * 64;
  d.gsg = (d.nbt + 1) * (d.nbt + 2) / 2 * 64;
  return d;
}

// extra shared-memory doubles behind the matrix region
__host__ __device__ inline int extra_doubles(const Dims& d) { return 8 * d.nbt + d.n2p + 2 * (d.nS + 22) + 64 + 8; }

// block-packed factor of the free block of the box QP (up to nx variables)
__host__ __device__ inline int qp_scratch_doubles(const Dims& d) {
  const int nbq = ((d.nx + 7) >> 3) + 1;  // + the right-hand-side block row
  return nbq * (nbq + 1) / 2 * 64;
}

__host__ __device__ inline int region_doubles(const Dims& d) {
  int r = d.gs;
  if (d.ga > r) r = d.ga;
  const int win = (d.m + d.p) * d.N + d.ell * (d.m + d.p) * 8 * ((d.m + d.p + 8) / 8);
  if (win > r) r = win;
  // box-QP stage: S7 | Hq | 20 vectors ; the block-packed factor scratch aliases S7
  const int qp = qp_scratch_doubles(d) + d.nx * d.ldq + 20 * kMaxNx;
  if (qp > r) r = qp;
  const int nr2 = 2 * d.nb3 + 1, zmat = nr2 * (nr2 + 1) / 2 * 64;  // [R3 | aug | E] matrix of the Z computation
  if (zmat > r) r = zmat;
  return (r + 1) & ~1;  // keep everything behind the region 16-byte aligned
}

__host__ __device__ inline size_t slot_doubles(const Dims& d) { return (size_t)d.ga + d.gsg; }

// ---- per-controller Gram cache (optional) ---------------------------------------------------------
// Between two solves of a closed loop the window moves by one sample, so the Hankel matrix drops its first column and
// gains a last one: G' = G - a a' + b b' with a = old first column, b = new last column.  With a cache the permuted,
// block-packed G of every controller persists in HBM (1.2 GB for 4096 default-size controllers) and a solve whose
// window moved by exactly one sample since the cached G applies the rank-2 correction block by block (coalesced 16-byte
// traffic, two DFMAs per entry) instead of regenerating G (DMMA first block column + 32 k scattered stores).  G is
// regenerated from the window when the cache is empty, the window moved by more than one sample, or `gram_refresh`
// incremental updates have accumulated (bounds the rounding drift: each update rounds at ulp(G_ij), the same size as the
// error of a fresh accumulation).  Per controller: kCacheHdr doubles of header
//   [0] int64 state: 0 empty, 1 valid, 2 valid + window moved one sample, 3 valid + moved more than one
//   [1] int64 age:   incremental updates since the last regeneration
//   [2 ..] the sample (u, y) that the last shift dropped
// followed by slot_doubles(d) doubles in the layout of a workspace slot.
constexpr int kCacheHdr = 24;
constexpr int kGramRefreshDefault = 128;
__host__ __device__ inline size_t cache_stride(const Dims& d) { return (size_t)kCacheHdr + slot_doubles(d); }
// called by the window-shifting kernels before they overwrite the first sample (threads 0 .. max(m, p) - 1 take part)
__device__ __forceinline__ void cache_note_shift(double* __restrict__ cache, const Dims& d, int b, const double* ub,
                                                 const double* yb) {
  if (!cache) return;
  double* h = cache + (size_t)b * cache_stride(d);
  const long long st = *reinterpret_cast<long long*>(h);
  if (st == 0) return;
  if (threadIdx.x < d.m) h[2 + threadIdx.x] = ub[threadIdx.x];
  if (threadIdx.x < d.p) h[2 + d.m + threadIdx.x] = yb[threadIdx.x];
  if (threadIdx.x == 0) *reinterpret_cast<long long*>(h) = st == 1 ? 2 : 3;
}

// ---- 8x8 block layout ---------------------------------------------------------------------------
// element (g, c) of a block sits at g*8 + (c ^ 4*((g>>1)&1)): every DMMA operand fragment (8 rows
// x 4 consecutive columns) covers all 32 banks exactly twice and every accumulator fragment is a
// contiguous 512-byte read of 16-byte pieces.
__device__ __forceinline__ int eoff(int g, int c) { return g * 8 + (c ^ ((g & 2) << 1)); }
__device__ __forceinline__ int tri(int I) { return I * (I + 1) / 2; }

template <bool RECT>
__device__ __forceinline__ int bidx(int I, int J, int nb1) {
  return RECT ? I * nb1 + J : tri(I) + J;
}

__device__ __forceinline__ double ld_ab(const double* blk, int g, int t, int h) {
  return blk[g * 8 + ((4 * h + t) ^ ((g & 2) << 1))];
}
__device__ __forceinline__ double2 ld_c(const double* blk, int g, int t) {
  return *reinterpret_cast<const double2*>(blk + g * 8 + ((2 * t) ^ ((g & 2) << 1)));
}
__device__ __forceinline__ void st_c(double* blk, int g, int t, double2 v) {
  *reinterpret_cast<double2*>(blk + g * 8 + ((2 * t) ^ ((g & 2) << 1))) = v;
}
// D(8x8) += A(8x4, row) * B(4x8, col): FP64 tensor core
__device__ __forceinline__ void dmma(double2& c, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c.x), "+d"(c.y)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// sum_{J < Jn} L_IJ L_KJ^T for up to two row blocks I0, I1 sharing the K operand
template <bool RECT>
__device__ __forceinline__ void syrk_pair(const double* __restrict__ M, int nb1, int I0, int I1, bool two, int K, int Jn,
                                          int g, int t, double2& s0, double2& s1) {
  double2 a0 = {0.0, 0.0}, a1 = {0.0, 0.0}, b0 = {0.0, 0.0}, b1 = {0.0, 0.0};
  const double* rK = M + (size_t)bidx<RECT>(K, 0, nb1) * 64;
  const double* r0 = M + (size_t)bidx<RECT>(I0, 0, nb1) * 64;
  const double* r1 = M + (size_t)bidx<RECT>(two ? I1 : I0, 0, nb1) * 64;
  for (int J = 0; J < Jn; ++J) {
    const double kb0 = ld_ab(rK + J * 64, g, t, 0), kb1 = ld_ab(rK + J * 64, g, t, 1);
    const double x0 = ld_ab(r0 + J * 64, g, t, 0), x1 = ld_ab(r0 + J * 64, g, t, 1);
    dmma(a0, x0, kb0);
    dmma(a1, x1, kb1);
    if (two) {
      const double y0 = ld_ab(r1 + J * 64, g, t, 0), y1 = ld_ab(r1 + J * 64, g, t, 1);
      dmma(b0, y0, kb0);
      dmma(b1, y1, kb1);
    }
  }
  s0.x = a0.x + a1.x; s0.y = a0.y + a1.y;
  s1.x = b0.x + b1.x; s1.y = b0.y + b1.y;
}

#define LT(i, c) a[(i) * ((i) + 1) / 2 + (c)]
// load the factor of a diagonal block (36 values, broadcast reads) into registers
__device__ __forceinline__ void load_diag(const double* dk, double (&a)[36]) {
  // 16-byte broadcast reads (the XOR-4 swizzle keeps even/odd column pairs adjacent)
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c <= i; c += 2) {
      const double2 v = *reinterpret_cast<const double2*>(dk + eoff(i, c));
      LT(i, c) = v.x;
      if (c + 1 <= i) LT(i, c + 1) = v.y;
    }
}

// Cholesky of one 8x8 diagonal block AND the inverse of its factor, every lane working on the whole block
// redundantly in registers: no shuffles and no shared-memory traffic on the 8-pivot dependency chain
// (rsqrt -> scale -> fma), no divergent code.  The factor is then inverted in place, column by column (column j
// of L is not needed once column j of L^-1 is done).  On exit the block holds W = L^-1 (zeros above the
// diagonal): every triangular solve against this block is then a product with W - two DMMAs for an 8x8 block
// (`trsm8_dmma`), an 8x8 mat-vec in the substitutions - instead of a dependent chain.  Published by identical
// 16-byte stores from all lanes.  (Interleaving the inverse with the pivots - the row operations of a forward
// substitution on the identity - was measured 40 % slower: the extra FP64 issues delay the pivot chain.)
#ifndef DDQC_INV_SPLIT
#define DDQC_INV_SPLIT 1  // 1: each lane inverts one column of the factor (0: every lane the whole factor)
#endif
__device__ __forceinline__ bool chol8_inv(double* blk, int lane) {
  double a[36];
  load_diag(blk, a);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double dj = LT(j, j);
    ok = ok && (dj > 0.0);
    dj = dj > 0.0 ? dj : 1.0;
    const double r = rsqrt(dj);
    LT(j, j) = r;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) LT(i, j) *= r;
#pragma unroll
    for (int c = j + 1; c < 8; ++c)
#pragma unroll
      for (int i = c; i < 8; ++i) LT(i, c) = fma(-LT(i, j), LT(c, j), LT(i, c));
  }
#if DDQC_INV_SPLIT
  // lane l computes column l & 7 of W: forward substitution of the unit vector e_j (static indices; the entries
  // above the diagonal come out as exact zeros)
  const int jc = lane & 7;
  double w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double s = i == jc ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-LT(i, k), w[k], s);
    w[i] = s * LT(i, i);
  }
  __syncwarp();
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) blk[eoff(i, jc)] = w[i];
  }
#else
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int i = j + 1; i < 8; ++i) {
      double s = LT(i, j) * LT(j, j);
#pragma unroll
      for (int k = j + 1; k < i; ++k) s = fma(LT(i, k), LT(k, j), s);
      LT(i, j) = -s * LT(i, i);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; c += 2)
      *reinterpret_cast<double2*>(blk + eoff(i, c)) = make_double2(c <= i ? LT(i, c) : 0.0, c + 1 <= i ? LT(i, c + 1) : 0.0);
#endif
  return ok;
}

// X <- X L^-T for one 8x8 block on the tensor cores: two DMMAs against W = L^-1 (B operand fragments kb0, kb1 of
// the diagonal block).  The block is read as an A operand and written back as an accumulator fragment; mma.sync
// orders the reads of all lanes before the writes.
__device__ __forceinline__ void trsm8_dmma(double* blk, double kb0, double kb1, int g, int t) {
  const double x0 = ld_ab(blk, g, t, 0), x1 = ld_ab(blk, g, t, 1);
  double2 acc = {0.0, 0.0};
  dmma(acc, x0, kb0);
  dmma(acc, x1, kb1);
  __syncwarp();
  st_c(blk, g, t, acc);
}

// Left-looking blocked Cholesky of block columns [0, ncols) of a matrix of `nrows` block rows held
// in shared memory (RECT: nrows x nb1 rectangle; else block-packed lower triangle).  Block rows
// >= ncols are carried along (they receive the triangular solve = forward substitution).
//
// The dependency chain of a Cholesky factorisation runs through the diagonal blocks:
//   factor (K,K) -> solve block (K+1,K) -> rank-8 update of (K+1,K+1) -> factor (K+1,K+1) -> ...
// Warp 0 owns exactly that chain and never waits for the bulk of the work; warps 1..15 own
// everything else and run one barrier behind it.  Per block column K:
//   warp 0 : factor (K,K), leaving W = L_KK^-1 in its place | barrier A | solve (K+1,K) | arrive B
//            | update (K+1,K+1) with column K
//   bulk   : update column K+1 with columns < K (the tensor-core bulk, overlaps warp 0's factor)
//            | barrier A | solve (I,K), I >= K+2 | wait B | update (I,K+1), I >= K+2 with column K
// Barrier A is a full 512-thread barrier (publishes W); B is asymmetric: warp 0 only arrives (it has nothing to
// wait for), the bulk warps synchronise on it.  Every solve is X <- X W^T on the tensor cores (trsm8_dmma).
// Measured and dropped (tools/ubench_chol8.cu, the DDQC_PROF_FACTOR stage counters): the FP64 pipe is per SM
// sub-partition and a dependent DFMA chain queued behind streaming DMMAs of the same sub-partition runs 5x slower,
// so the chain of warp 0 (~2.3 k cycles per column, 1.1 k alone) and the bulk (~2.5 k) are within 10 % of each
// other; taking warps 4, 8, 12 out of the lookahead (DDQC_BULK_ALL=0) moves the bound from the chain to the bulk
// without changing the total, and factoring block columns in pairs (one barrier pair per 16 columns, six
// independent DMMAs per block row against the inverse of the 16x16 diagonal factor, a helper warp for the second
// chain row) cut the lookahead by a third but lengthened the chain by as much.
__device__ __forceinline__ void bar_sync_named(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kThreads) : "memory"); }

// developer sub-phase probe (-DDDQC_PROF_SUB; slots 24..39 of the phase buffer): a barrier + the time since the
// previous probe point.  It perturbs the schedule and is compiled out of the product.
#ifdef DDQC_PROF_SUB
__device__ long long* g_subclk = nullptr;
__shared__ long long s_tsub;
#define DDQC_SUB(k)                                                                                          \
  do {                                                                                                       \
    __syncthreads();                                                                                         \
    if (g_subclk && threadIdx.x == 0) {                                                                      \
      const long long t_now = clock64();                                                                     \
      if ((k) >= 0) atomicAdd((unsigned long long*)&g_subclk[24 + (k)], (unsigned long long)(t_now - s_tsub)); \
      s_tsub = t_now;                                                                                        \
    }                                                                                                        \
  } while (0)
#else
#define DDQC_SUB(k)
#endif

#ifdef DDQC_PROF_FACTOR
__device__ long long* g_fprof = nullptr;
#define FPROF(k)                                                                            \
  do {                                                                                      \
    if (g_fprof && lane == 0 && (warp == 0 || warp == 2)) {                                                 \
      const long long tn = clock64();                                                       \
      atomicAdd((unsigned long long*)&g_fprof[12 + (warp ? 6 : 0) + (k)], (unsigned long long)(tn - tp)); \
      tp = tn;                                                                              \
    }                                                                                       \
  } while (0)
#else
#define FPROF(k)
#endif

// The solver inlines to 441 KB of SASS and ncu shows an instruction-cache hit rate of 83 % (stall "no instruction" 0.5
// per issue).  DDQC_MPC_OUTLINE = 1 keeps ONE body of each of the large phase routines (called 2-4 times each; 366 KB).
// Measured and left off: 514-518 k solves/s against 614 k on the phase probe (tools/mpc_perf.py) - behind a call the
// routines lose the constant propagation of their shape arguments and the pivoting pays the ABI spills (96 -> 145 k
// cycles per solve), far more than the instruction fetch gains.
#ifndef DDQC_MPC_OUTLINE
#define DDQC_MPC_OUTLINE 0
#endif
#if DDQC_MPC_OUTLINE
#define DDQC_PHASE_FN __noinline__
#else
#define DDQC_PHASE_FN
#endif
template <bool RECT>
__device__ DDQC_PHASE_FN void factor_columns(double* __restrict__ M, int nb1, int nrows, int ncols, int* s_flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  constexpr int kBulk = kWarps - 1;
  __syncthreads();
#ifdef DDQC_PROF_FACTOR
  long long tp = clock64();
#endif
  for (int K = 0; K < ncols; ++K) {
    double* dk = M + (size_t)bidx<RECT>(K, K, nb1) * 64;
    if (warp == 0) {
      if (!chol8_inv(dk, lane) && lane == 0) *s_flag = 1;  // the diagonal block now holds W = L_KK^-1
      __syncwarp();
      FPROF(0);
      bar_sync_named(1);
      FPROF(1);
      if (K + 1 < nrows) {
        trsm8_dmma(M + (size_t)bidx<RECT>(K + 1, K, nb1) * 64, ld_ab(dk, g, t, 0), ld_ab(dk, g, t, 1), g, t);
        __syncwarp();
      }
      FPROF(2);
      bar_arrive_named(2);
      if (K + 1 < ncols) {
        const double* rk = M + (size_t)bidx<RECT>(K + 1, K, nb1) * 64;
        double* dn = M + (size_t)bidx<RECT>(K + 1, K + 1, nb1) * 64;
        const double x0 = ld_ab(rk, g, t, 0), x1 = ld_ab(rk, g, t, 1);
        double2 acc = {0.0, 0.0}, cv = ld_c(dn, g, t);
        dmma(acc, x0, x0);
        dmma(acc, x1, x1);
        cv.x -= acc.x; cv.y -= acc.y;
        st_c(dn, g, t, cv);
        __syncwarp();
      }
      FPROF(3);
    } else {
      const int bw = warp - 1;
      // DDQC_BULK_ALL == 0: warps 4, 8, 12 share warp 0's SM sub-partition and its FP64 pipe and sit out the
      // tensor-core bulk so that the diagonal-block chain is not queued behind 16-cycle DMMA issues
      if (K > 0 && K + 1 < ncols && (DDQC_BULK_ALL || (warp & 3) != 0)) {
        const int bw12 = DDQC_BULK_ALL ? bw : (warp >> 2) * 3 + (warp & 3) - 1;
        for (int I = K + 1 + 2 * bw12; I < nrows; I += 2 * (DDQC_BULK_ALL ? kBulk : kWarps - kWarps / 4)) {
          const bool two = I + 1 < nrows;
          double2 s0, s1;
          syrk_pair<RECT>(M, nb1, I, I + 1, two, K + 1, K, g, t, s0, s1);
          double* b0 = M + (size_t)bidx<RECT>(I, K + 1, nb1) * 64;
          double2 cv = ld_c(b0, g, t);
          cv.x -= s0.x; cv.y -= s0.y;
          st_c(b0, g, t, cv);
          if (two) {
            double* b1 = M + (size_t)bidx<RECT>(I + 1, K + 1, nb1) * 64;
            double2 cw = ld_c(b1, g, t);
            cw.x -= s1.x; cw.y -= s1.y;
            st_c(b1, g, t, cw);
          }
        }
      }
      FPROF(0);
      bar_sync_named(1);
      FPROF(1);
      if (K + 2 + bw < nrows) {
        const double ib0 = ld_ab(dk, g, t, 0), ib1 = ld_ab(dk, g, t, 1);
        for (int I = K + 2 + bw; I < nrows; I += kBulk) trsm8_dmma(M + (size_t)bidx<RECT>(I, K, nb1) * 64, ib0, ib1, g, t);
      }
      FPROF(2);
      bar_sync_named(2);
      FPROF(3);
      if (K + 1 < ncols && K + 2 + bw < nrows) {
        const double* rk = M + (size_t)bidx<RECT>(K + 1, K, nb1) * 64;
        const double kb0 = ld_ab(rk, g, t, 0), kb1 = ld_ab(rk, g, t, 1);
        for (int I = K + 2 + bw; I < nrows; I += kBulk) {
          const double* li = M + (size_t)bidx<RECT>(I, K, nb1) * 64;
          double* b0 = M + (size_t)bidx<RECT>(I, K + 1, nb1) * 64;
          double2 acc = {0.0, 0.0}, cv = ld_c(b0, g, t);
          dmma(acc, ld_ab(li, g, t, 0), kb0);
          dmma(acc, ld_ab(li, g, t, 1), kb1);
          cv.x -= acc.x; cv.y -= acc.y;
          st_c(b0, g, t, cv);
        }
      }
      FPROF(4);
    }
  }
  __syncthreads();
}

// y_K <- W y_K (TRANS: W^T y_K) for one diagonal block holding W = L_KK^-1, by lanes 0..7 of one warp
template <bool TRANS>
__device__ __forceinline__ void diag_apply(const double* __restrict__ dk, double* __restrict__ yk, int lane) {
  double acc = 0.0;
  if (lane < 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) acc = fma(TRANS ? dk[eoff(c, lane)] : dk[eoff(lane, c)], yk[c], acc);
  }
  __syncwarp();
  if (lane < 8) yk[lane] = acc;
  __syncwarp();
}

// L y' = y (in place) for the leading ncols block columns of a factored block-packed matrix (diagonal blocks hold
// the inverses of their factors).  One barrier per block column: warp 0 takes block row K+1 of the update and
// applies the next diagonal inverse right away, the other warps update the rows below.
__device__ void fwdsub_blocked(const double* __restrict__ M, int ncols, double* __restrict__ y) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0 && ncols > 0) diag_apply<false>(M, y, lane);
  __syncthreads();
  for (int K = 0; K + 1 < ncols; ++K) {
    if (warp == 0) {
      const double* blk = M + (size_t)(tri(K + 1) + K) * 64;
      if (lane < 8) {
        double acc = y[8 * (K + 1) + lane];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = fma(-blk[eoff(lane, c)], y[8 * K + c], acc);
        y[8 * (K + 1) + lane] = acc;
      }
      __syncwarp();
      diag_apply<false>(M + (size_t)(tri(K + 1) + K + 1) * 64, y + 8 * (K + 1), lane);
    } else {
      for (int e = tid - 32; e < 8 * (ncols - K - 2); e += kThreads - 32) {
        const int I = K + 2 + (e >> 3), gr = e & 7;
        const double* blk = M + (size_t)(tri(I) + K) * 64;
        double acc = y[8 * I + gr];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = fma(-blk[eoff(gr, c)], y[8 * K + c], acc);
        y[8 * I + gr] = acc;
      }
    }
    __syncthreads();
  }
}

// L^T z = y (in place), same organisation from the last block column upwards
__device__ void backsub_blocked(const double* __restrict__ M, int ncols, double* __restrict__ y) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0 && ncols > 0) diag_apply<true>(M + (size_t)(tri(ncols - 1) + ncols - 1) * 64, y + 8 * (ncols - 1), lane);
  __syncthreads();
  for (int K = ncols - 1; K > 0; --K) {
    if (warp == 0) {
      const double* blk = M + (size_t)(tri(K) + K - 1) * 64;
      if (lane < 8) {
        double acc = y[8 * (K - 1) + lane];
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) acc = fma(-blk[eoff(gg, lane)], y[8 * K + gg], acc);
        y[8 * (K - 1) + lane] = acc;
      }
      __syncwarp();
      diag_apply<true>(M + (size_t)(tri(K - 1) + K - 1) * 64, y + 8 * (K - 1), lane);
    } else {
      for (int e = tid - 32; e < 8 * (K - 1); e += kThreads - 32) {
        const int J = e >> 3, c = e & 7;
        const double* blk = M + (size_t)(tri(K) + J) * 64;
        double acc = y[e];
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) acc = fma(-blk[eoff(gg, c)], y[8 * K + gg], acc);
        y[e] = acc;
      }
    }
    __syncthreads();
  }
}

// Schur complement of the trailing block columns [c0, nrows): C_IK -= sum_{J < c0} L_IJ L_KJ^T
__device__ DDQC_PHASE_FN void schur_trailing(double* __restrict__ M, int nrows, int c0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int nrem = nrows - c0;
  const int nblk = nrem * (nrem + 1) / 2;
  for (int b = warp; b < nblk; b += kWarps) {
    int i = (int)((sqrtf(8.0f * b + 1.0f) - 1.0f) * 0.5f);
    while (tri(i + 1) <= b) ++i;
    while (tri(i) > b) --i;
    const int j = b - tri(i);
    double2 s0, s1;
    syrk_pair<false>(M, 0, c0 + i, 0, false, c0 + j, c0, g, t, s0, s1);
    double* blk = M + (size_t)(tri(c0 + i) + c0 + j) * 64;
    double2 cv = ld_c(blk, g, t);
    cv.x -= s0.x; cv.y -= s0.y;
    st_c(blk, g, t, cv);
  }
}

// workspace slot (global, L2-resident) -> shared memory, n2 16-byte pieces, as asynchronous copies that do not
// pass through registers: all of a thread's pieces are in flight at once (the plain load/store loop waits for an
// L2 round trip every four pieces).  Ends with the data visible to the calling thread; the caller synchronises.
__device__ __forceinline__ void copy_slot_to_smem(double* __restrict__ dst, const double* __restrict__ src, int n2) {
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(dst);
  const double2* s2 = reinterpret_cast<const double2*>(src);
  for (int e = threadIdx.x; e < n2; e += kThreads)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + 16u * (uint32_t)e), "l"(s2 + e) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// pointer to element (hi, lo), hi >= lo, of the block-packed lower-triangular matrix
__device__ __forceinline__ double* elem_tri(double* M, int hi, int lo) {
  return M + (size_t)(tri(hi >> 3) + (lo >> 3)) * 64 + eoff(hi & 7, lo & 7);
}

// ---- small dense helpers for the box QP (whole CTA, uniform control flow) --------------------------

// out = Hq x + f, 4 threads per row
__device__ void matvec_sym(const double* __restrict__ Hq, int ld, int nx, const double* __restrict__ x,
                           const double* __restrict__ f, double* __restrict__ out) {
  const int tid = threadIdx.x, part = tid & 3;
  for (int r0 = 0; r0 < nx; r0 += kThreads / 4) {
    const int row = r0 + (tid >> 2);
    double acc = 0.0;
    if (row < nx)
      for (int j = part; j < nx; j += 4) acc += Hq[row * ld + j] * x[j];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (row < nx && part == 0) out[row] = acc + f[row];
  }
}

enum { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };
// block-wide reduction; every thread passes a value and receives the result (two barriers)
__device__ double block_reduce(double v, int op, double* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == RED_SUM ? v + w : (op == RED_MAX ? fmax(v, w) : fmin(v, w));
  }
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double r = s_red[0];
  for (int w = 1; w < kWarps; ++w) r = op == RED_SUM ? r + s_red[w] : (op == RED_MAX ? fmax(r, s_red[w]) : fmin(r, s_red[w]));
  return r;
}

struct QpWork {
  const double* Hq; int ldq; int nx;
  double* F; int ldf;          // factor scratch (>= nx x nx)
  double *f, *lo, *hi, *x, *g, *rhs, *rs;
  double *sl, *su, *zl, *zu, *dx, *dzl, *dzu, *rd, *dg;
  double* s_red;
  int *state, *idx, *verdict, *s_flag;
};

// Gather the principal submatrix Hq[idx, idx] (idx = q.idx when `use_idx`, identity otherwise) plus an
// optional diagonal into block-packed storage at q.F (padded with identity rows to a multiple of 8)
// and factor it on the tensor cores.  With `rhs` the right-hand side rides along as row 0 of an extra block
// row: the triangular solves of the factorisation forward-substitute it for free (y = L^-1 rhs ends up in
// that row), so only the back-substitution remains.
__device__ void qp_factor(const QpWork& q, int k, bool use_idx, const double* diag, const double* rhs) {
  const int tid = threadIdx.x;
  const int nbq = (k + 7) >> 3, kp = nbq * 8, nrows = nbq + (rhs ? 1 : 0);
  double2* z = reinterpret_cast<double2*>(q.F);
  for (int e = tid; e < tri(nrows) * 32; e += kThreads) z[e] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int a = tid >> 5; a < kp + (rhs ? 1 : 0); a += kWarps) {
    const int ia = a < k ? (use_idx ? q.idx[a] : a) : 0;
    for (int c = tid & 31; c <= a && c < kp; c += 32) {
      double v;
      if (a < k) {
        const int ic = use_idx ? q.idx[c] : c;
        v = q.Hq[ia * q.ldq + ic];
        if (a == c && diag) v += diag[a];
      } else if (a < kp) {
        v = a == c ? 1.0 : 0.0;
      } else {
        v = c < k ? rhs[c] : 0.0;
      }
      *elem_tri(q.F, a, c) = v;
    }
  }
  __syncthreads();
  DDQC_SUB(13);
  factor_columns<false>(q.F, 0, nrows, nbq, q.s_flag);
}

// One pass of block principal pivoting in the free space.  Returns the number of infeasible
// indices (0 = optimal) and the largest infeasible index through `last`.
__device__ int bpp_pass(const QpWork& q, double tol_g, int* last) {
  const int tid = threadIdx.x, lane = tid & 31, nx = q.nx;
  // x at the bounds of the active set, free list
  if (tid < nx) {
    const int st = q.state[tid];
    q.x[tid] = st < 0 ? q.lo[tid] : (st > 0 ? q.hi[tid] : 0.0);
  }
  if (tid < 32) {
    int nf = 0;
    for (int base = 0; base < nx; base += 32) {
      const int i = base + lane;
      const bool fr = i < nx && q.state[i] == 0;
      const unsigned mk = __ballot_sync(0xffffffffu, fr);
      if (fr) q.idx[nf + __popc(mk & ((1u << lane) - 1u))] = i;
      nf += __popc(mk);
    }
    if (lane == 0) q.idx[kMaxNx] = nf;
  }
  __syncthreads();
  const int kf = q.idx[kMaxNx];
  DDQC_SUB(-1);
  matvec_sym(q.Hq, q.ldq, nx, q.x, q.f, q.g);
  __syncthreads();
  DDQC_SUB(8);
  const int nbf = (kf + 7) >> 3;
  if (tid < kf) q.rhs[tid] = -q.g[q.idx[tid]];
  __syncthreads();
  qp_factor(q, kf, true, nullptr, q.rhs);
  DDQC_SUB(9);
  if (tid < 8 * nbf) q.rhs[tid] = *elem_tri(q.F, 8 * nbf, tid);  // y = L^-1 rhs
  __syncthreads();
  backsub_blocked(q.F, nbf, q.rhs);
  DDQC_SUB(10);
  if (tid < kf) q.x[q.idx[tid]] = q.rhs[tid];
  __syncthreads();
  matvec_sym(q.Hq, q.ldq, nx, q.x, q.f, q.g);
  __syncthreads();
  DDQC_SUB(11);
  const double tol_x = 1e-10;
  int v = 0;
  if (tid < nx) {
    const int st = q.state[tid];
    const double gi = q.g[tid], xi = q.x[tid];
    if (st < 0 && gi < -tol_g) v = 2;
    else if (st > 0 && gi > tol_g) v = 2;
    else if (st == 0) {
      if (xi < q.lo[tid] - tol_x * (1.0 + fabs(q.lo[tid]))) v = -1;
      else if (xi > q.hi[tid] + tol_x * (1.0 + fabs(q.hi[tid]))) v = 1;
    }
    q.verdict[tid] = v;
  }
  const int ninf = __syncthreads_count(v != 0);
  const double lastd = block_reduce(v != 0 ? (double)tid : -1.0, RED_MAX, q.s_red);
  *last = (int)lastd;
  DDQC_SUB(12);
  return ninf;
}

// pivoting driver; returns 1 when converged.  `passes` accumulates the number of passes.
__device__ DDQC_PHASE_FN int bpp_solve(const QpWork& q, double tol_g, int max_pass, int* passes) {
  const int tid = threadIdx.x, nx = q.nx;
  int best = nx + 1, patience = 3;
  for (int it = 0; it < max_pass; ++it) {
    int last;
    const int ninf = bpp_pass(q, tol_g, &last);
    ++*passes;
    if (ninf == 0) return 1;
    bool full;
    if (ninf < best) { best = ninf; patience = 3; full = true; }
    else if (patience > 0) { --patience; full = true; }
    else full = false;
    if (tid < nx && (full || tid == last)) {
      const int v = q.verdict[tid];
      if (v == -1) q.state[tid] = -1;
      else if (v == 1) q.state[tid] = 1;
      else if (v == 2) q.state[tid] = 0;
    }
    __syncthreads();
  }
  return 0;
}

// Mehrotra predictor-corrector interior point method for the box QP.  Leaves x in q.x and an
// active-set estimate in q.state.  Returns iterations + 1 (negated when the cap was reached).
__device__ int ipm_solve(const QpWork& q, int max_it) {
  const int tid = threadIdx.x, lane = tid & 31, nx = q.nx;
  const bool act = tid < nx;
  const bool il = act && isfinite(q.lo[act ? tid : 0]);
  const bool iu = act && isfinite(q.hi[act ? tid : 0]);
  const double lo = il ? q.lo[tid] : 0.0, hi = iu ? q.hi[tid] : 0.0;
  double x = 0.0, sl = 1.0, su = 1.0, zl = 0.0, zu = 0.0;
  if (act) {
    x = (il && iu) ? 0.5 * (lo + hi) : (il ? lo + 1.0 : (iu ? hi - 1.0 : 0.0));
    if (il) sl = x - lo;
    if (iu) su = hi - x;
    q.x[tid] = x;
  }
  __syncthreads();
  matvec_sym(q.Hq, q.ldq, nx, q.x, q.f, q.g);
  __syncthreads();
  const double sc = fmax(1.0, block_reduce(act ? fabs(q.g[tid]) : 0.0, RED_MAX, q.s_red));
  if (il) zl = sc;
  if (iu) zu = sc;
  const double nc = block_reduce((il ? 1.0 : 0.0) + (iu ? 1.0 : 0.0), RED_SUM, q.s_red);
  int it = 0;
  bool done = false;
  for (; it < max_it; ++it) {
    if (act) q.x[tid] = x;
    __syncthreads();
    matvec_sym(q.Hq, q.ldq, nx, q.x, q.f, q.g);
    __syncthreads();
    const double gi = act ? q.g[tid] : 0.0;
    const double rd = gi - zl + zu;
    const double mu = block_reduce((il ? sl * zl : 0.0) + (iu ? su * zu : 0.0), RED_SUM, q.s_red) / nc;
    const double rdmax = block_reduce(fabs(rd), RED_MAX, q.s_red);
    if (rdmax < 1e-9 * sc && mu < 1e-12 * sc) { done = true; break; }
    // F = Hq + diag(zl/sl + zu/su)
    const double dg = (il ? zl / sl : 0.0) + (iu ? zu / su : 0.0);
    if (act) q.dg[tid] = dg;
    __syncthreads();
    qp_factor(q, nx, false, q.dg, nullptr);
    const int nbq = (nx + 7) >> 3;
    if (tid < 8 * nbq) q.rhs[tid] = act ? -gi : 0.0;  // affine right-hand side
    __syncthreads();
    fwdsub_blocked(q.F, nbq, q.rhs);
    backsub_blocked(q.F, nbq, q.rhs);
    const double dxa = act ? q.rhs[tid] : 0.0;
    const double dzla = il ? (-sl * zl - zl * dxa) / sl : 0.0;
    const double dzua = iu ? (-su * zu + zu * dxa) / su : 0.0;
    double a1 = 1.0;
    if (il && dxa < 0.0) a1 = fmin(a1, -sl / dxa);
    if (iu && dxa > 0.0) a1 = fmin(a1, su / dxa);
    if (il && dzla < 0.0) a1 = fmin(a1, -zl / dzla);
    if (iu && dzua < 0.0) a1 = fmin(a1, -zu / dzua);
    const double aa = block_reduce(a1, RED_MIN, q.s_red);
    const double mu_aff =
        block_reduce((il ? (sl + aa * dxa) * (zl + aa * dzla) : 0.0) + (iu ? (su - aa * dxa) * (zu + aa * dzua) : 0.0),
                     RED_SUM, q.s_red) / nc;
    const double rat = mu_aff / mu;
    const double sig = rat * rat * rat;
    const double rcl = sig * mu - dxa * dzla, rcu = sig * mu + dxa * dzua;
    if (tid < 8 * nbq) q.rhs[tid] = act ? -rd + (il ? (rcl - sl * zl) / sl : 0.0) - (iu ? (rcu - su * zu) / su : 0.0) : 0.0;
    __syncthreads();
    fwdsub_blocked(q.F, nbq, q.rhs);
    backsub_blocked(q.F, nbq, q.rhs);
    const double dx = act ? q.rhs[tid] : 0.0;
    const double dzl = il ? (rcl - sl * zl - zl * dx) / sl : 0.0;
    const double dzu = iu ? (rcu - su * zu + zu * dx) / su : 0.0;
    double a2 = 1.0;
    if (il && dx < 0.0) a2 = fmin(a2, -sl / dx);
    if (iu && dx > 0.0) a2 = fmin(a2, su / dx);
    if (il && dzl < 0.0) a2 = fmin(a2, -zl / dzl);
    if (iu && dzu < 0.0) a2 = fmin(a2, -zu / dzu);
    double al = block_reduce(a2, RED_MIN, q.s_red);
    if (al < 1.0) al *= 0.995;
    x += al * dx;
    if (il) { sl += al * dx; zl += al * dzl; }
    if (iu) { su -= al * dx; zu += al * dzu; }
  }
  if (act) {
    q.x[tid] = x;
    q.state[tid] = (il && zl > sl) ? -1 : ((iu && zu > su) ? 1 : 0);
  }
  __syncthreads();
  return done ? (it + 1) : -(it + 1);
}

// ---- two staging steps of the Z computation, kept out of line so that their register arrays do not weigh on
// the register allocation of the rest of the solver --------------------------------------------------------

// S (lower blocks (nb2+i, nb2+j) of the factored matrix) -> [S | E] from offset 0, E = identity below S
__device__ __noinline__ void repack_schur(double* __restrict__ sM, int nb2, int nb3) {
  const int tid = threadIdx.x;
  const int nsb = nb3 + 1, nsrc = tri(nsb) * 64, nr2 = 2 * nb3 + 1;
  double keep[kRepackQ];  // source and destination overlap: through registers
#pragma unroll
  for (int qq = 0; qq < kRepackQ; ++qq) {
    const int e = tid + qq * kThreads;
    keep[qq] = 0.0;
    if (e < nsrc) {
      const int bb = e >> 6;
      int i = (int)((sqrtf(8.0f * bb + 1.0f) - 1.0f) * 0.5f);
      while (tri(i + 1) <= bb) ++i;
      while (tri(i) > bb) --i;
      const int j = bb - tri(i);
      keep[qq] = sM[(size_t)(tri(nb2 + i) + nb2 + j) * 64 + (e & 63)];
    }
  }
  __syncthreads();
  double2* z2 = reinterpret_cast<double2*>(sM);
  for (int e = tid; e < tri(nr2) * 32; e += kThreads) z2[e] = make_double2(0.0, 0.0);
  __syncthreads();
#pragma unroll
  for (int qq = 0; qq < kRepackQ; ++qq) {
    const int e = tid + qq * kThreads;
    if (e < nsrc) sM[e] = keep[qq];  // same block order (tri(i) + j), now from offset 0
  }
  for (int e = tid; e < 8 * nb3; e += kThreads) *elem_tri(sM, 8 * nsb + e, e) = 1.0;
  __syncthreads();
}

// Hq = scale * Z[:nx, :nx], f = scale * Z[nx, :nx] read from the trailing blocks of the [R3 | aug | E] matrix;
// Hq overlaps the matrix it is read from: through registers.  A warp takes whole rows, lanes the columns.
__device__ __noinline__ void assemble_hq(const double* __restrict__ sM, double* __restrict__ Hq, double* __restrict__ f,
                                         double scale, int nb3, int n3, int nx, int ldq) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pa = 8 * nb3, pe = 8 * (nb3 + 1);
  double keep[kAsmRows][kAsmCols];
#pragma unroll
  for (int rr = 0; rr < kAsmRows; ++rr) {
    const int i = warp + rr * kWarps;  // i in [0, nx]: row nx is the c-row (gradient)
    const int ii = i < nx ? i : nx;
    const int Pi = ii < n3 ? pe + ii : pa + (ii - n3);
#pragma unroll
    for (int kk = 0; kk < kAsmCols; ++kk) {
      const int k = lane + 32 * kk;
      keep[rr][kk] = 0.0;
      if (i <= nx && k < nx) {
        const int Pk = k < n3 ? pe + k : pa + (k - n3);
        const int hi = Pi > Pk ? Pi : Pk, lo = Pi > Pk ? Pk : Pi;
        keep[rr][kk] = scale * sM[(size_t)(tri(hi >> 3) + (lo >> 3)) * 64 + eoff(hi & 7, lo & 7)];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < kAsmRows; ++rr) {
    const int i = warp + rr * kWarps;
#pragma unroll
    for (int kk = 0; kk < kAsmCols; ++kk) {
      const int k = lane + 32 * kk;
      if (i <= nx && k < nx) {
        if (i < nx) Hq[i * ldq + k] = keep[rr][kk];
        else f[k] = keep[rr][kk];
      }
    }
  }
  __syncthreads();
}

// ---- tiny box QP (<= 6 variables): min 0.5 x'Hx + f'x, lo <= x <= hi, H SPD -----------------------------------
// Used for the optimal reachable equilibrium (u_sr, y_sr) of the alpha_sr system.  The system of a given active set
// keeps the fixed size 6 (active variables and unused slots become identity rows with the bound as right-hand
// side), so every array index is static and everything lives in registers.  Inputs (H 36, f 6, lo 6, hi 6
// doubles) are read from shared memory at `io`.  Not inlined: their ~100 live registers must not weigh on the
// register allocation of the solver kernel.

// x for the active set st (0 free, -1 at lo, +1 at hi, 2 unused slot); false when a pivot is not positive.
// H, f, lo, hi stay in shared memory (io); only the lower triangle of the working matrix (LDL^T, 21 entries), the
// right-hand side and x live in registers, so that the routine fits the 128 registers of the solver kernel
// without spilling (the first version kept H and a full 6x6 copy: 1.7 kB of spill traffic per call).
#define A6(a, j) A[(a) * ((a) + 1) / 2 + (j)]
__device__ __forceinline__ bool qp6_solve_set(const double* __restrict__ io, const int (&st)[6], double (&x)[6]) {
  double A[21], r[6], xb[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) xb[i] = st[i] == 0 ? 0.0 : (st[i] == 2 ? 0.0 : (st[i] < 0 ? io[42 + i] : io[48 + i]));
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    double ra = -io[36 + a];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const bool fa = st[a] == 0, fj = st[j] == 0;
      const double h = io[a * 6 + j];
      if (j <= a) A6(a, j) = (fa && fj) ? h : (a == j ? 1.0 : 0.0);
      if (fa && !fj) ra -= h * xb[j];
    }
    r[a] = st[a] == 0 ? ra : xb[a];
  }
  // LDL^T of the SPD system (no pivoting needed); the diagonal keeps 1 / d
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    ok = ok && A6(c, c) > 0.0;
    const double inv = 1.0 / A6(c, c);
    A6(c, c) = inv;
    double tc[6];
#pragma unroll
    for (int a = c + 1; a < 6; ++a) tc[a] = A6(a, c);
#pragma unroll
    for (int a = c + 1; a < 6; ++a) {
      const double l = tc[a] * inv;
#pragma unroll
      for (int j = c + 1; j <= a; ++j) A6(a, j) -= l * tc[j];
      r[a] -= l * r[c];
      A6(a, c) = l;
    }
  }
#pragma unroll
  for (int a = 5; a >= 0; --a) {
    double v = r[a] * A6(a, a);
#pragma unroll
    for (int c = a + 1; c < 6; ++c) v -= A6(c, a) * x[c];
    x[a] = v;
  }
  return ok;
}

// gradient component i of the QP at x
__device__ __forceinline__ double qp6_grad(const double* __restrict__ io, int i, const double (&x)[6]) {
  double gi = io[36 + i];
#pragma unroll
  for (int j = 0; j < 6; ++j) gi += io[i * 6 + j] * x[j];
  return gi;
}

// Serial version: block principal pivoting by one thread.  Result at io[54..59].
__device__ __noinline__ bool small_box_qp(double* io, int nv) {
  double x[6];
  int st[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { st[i] = i < nv ? 0 : 2; x[i] = 0.0; }
  for (int pass = 0; pass < 64; ++pass) {
    if (!qp6_solve_set(io, st, x)) return false;
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (st[i] == 0) {
        if (x[i] < io[42 + i]) { st[i] = -1; changed = true; }
        else if (x[i] > io[48 + i]) { st[i] = 1; changed = true; }
      } else if (st[i] != 2) {
        const double gi = qp6_grad(io, i, x);
        if ((st[i] < 0 && gi < 0.0) || (st[i] > 0 && gi > 0.0)) { st[i] = 0; changed = true; }
      }
    }
    if (!changed) {
#pragma unroll
      for (int i = 0; i < 6; ++i) io[54 + i] = x[i];
      return true;
    }
  }
  return false;
}

// Parallel version: the calling thread tests ONE active set - digit i (base 3) of `combo` says whether the bounded
// variable i < nbnd is free, at its lower or at its upper bound; variables >= nbnd are free - and returns whether
// its solution satisfies the optimality conditions.  With nbnd <= 5 all 3^nbnd sets are tested at once by
// different threads (one solve instead of a pass per change of the active set); the QP is strictly convex, so
// exactly one set passes (several only when a variable sits on a bound with a zero multiplier - any of them is
// the optimum).
__device__ __noinline__ bool small_box_qp_combo(const double* io, int nv, int nbnd, int combo, double (&x)[6]) {
  int st[6];
  int cc = combo;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    st[i] = i < nv ? 0 : 2;
    if (i < nbnd) {
      const int dgt = cc % 3;
      cc /= 3;
      st[i] = dgt == 0 ? 0 : (dgt == 1 ? -1 : 1);
    }
    x[i] = 0.0;
  }
  bool ok = qp6_solve_set(io, st, x);
  double fs = 1.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) fs = fmax(fs, fabs(io[36 + i]));
  const double tol_g = 1e-9 * fs;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    if (st[i] == 0) {
      const double lo = io[42 + i], hi = io[48 + i];
      ok = ok && (x[i] >= lo - 1e-12 * (1.0 + fabs(lo))) && (x[i] <= hi + 1e-12 * (1.0 + fabs(hi)));
    } else if (st[i] != 2) {
      const double gi = qp6_grad(io, i, x);
      ok = ok && (st[i] < 0 ? gi >= -tol_g : gi <= tol_g);
    }
  }
  return ok;
}

// ---- Gram generation -------------------------------------------------------------------------------
// position of Hankel row (signal a, block row j) in the elimination order [R1 | R2 | R3 | aug]
__device__ __forceinline__ int pos_of(const Dims& d, int a, int j) {
  if (a < d.m) {
    if (j < d.n) return j * d.m + a;
    if (j >= d.L) return d.n * d.m + 1 + (j - d.L) * d.m + a;
    return d.n1p + d.n2p + (j - d.n) * d.m + a;
  }
  return d.n1p + j * d.p + (a - d.m);
}

// write element (P, Q) of the permuted, padded Gram matrix into the workspace slot (GA panel or
// GS trailing), mirroring inside diagonal blocks so that every block is fully defined
__device__ __forceinline__ void put_sym(int n1p, int nb1, double* __restrict__ GA, double* __restrict__ GS, int P, int Q,
                                        double v) {
  const int hi = P > Q ? P : Q, lo = P > Q ? Q : P;
  if (lo < n1p) {
    double* blk = GA + (size_t)((hi >> 3) * nb1 + (lo >> 3)) * 64;
    blk[eoff(hi & 7, lo & 7)] = v;
    if ((hi >> 3) == (lo >> 3)) blk[eoff(lo & 7, hi & 7)] = v;
  } else {
    const int h = hi - n1p, l = lo - n1p;
    double* blk = GS + (size_t)(tri(h >> 3) + (l >> 3)) * 64;
    blk[eoff(h & 7, l & 7)] = v;
    if ((h >> 3) == (l >> 3)) blk[eoff(l & 7, h & 7)] = v;
  }
}
__device__ __forceinline__ void put_sym(const Dims& d, double* __restrict__ GA, double* __restrict__ GS, int P, int Q,
                                        double v) {
  put_sym(d.n1p, d.nb1, GA, GS, P, Q, v);
}

// Every entry of G outside its first block column follows from the Hankel structure by a two-term recurrence along
// the block diagonals.  A task walks two diagonals whose lengths add up to ~ell, so all tasks cost the same:
// (a < b, u): delta = u and delta = u - ell;  (a == a, u): delta = u and ell-1-u.  Kept out of line: the loop runs
// ~32 k times per solve and must not inherit the register pressure (spills) of the solver kernel around it.
__device__ __noinline__ void gram_walk(int n1p, int nb1, int N, int ell, int C, int nsig, const double* __restrict__ sW,
                                       const double* __restrict__ sT, int ldT, const short* __restrict__ s_ppos,
                                       double* __restrict__ GA, double* __restrict__ GS) {
  const int nlt = nsig * (nsig - 1) / 2, half = (ell + 1) / 2;
  const int ntask = nlt * ell + nsig * half;
  for (int tk = threadIdx.x; tk < ntask; tk += kThreads) {
    int a, bs, u, d0, d1;
    if (tk < nlt * ell) {
      int pi = tk / ell;
      u = tk % ell;
      a = 0;
      while (pi >= nsig - 1 - a) { pi -= nsig - 1 - a; ++a; }
      bs = a + 1 + pi;
      d0 = u;
      d1 = u - ell;  // invalid (-ell) for u == 0
    } else {
      const int r2 = tk - nlt * ell;
      a = bs = r2 / half;
      u = r2 % half;
      d0 = u;
      d1 = ell - 1 - u;
      if (d1 == d0) d1 = -ell;
    }
    const double* wa = sW + a * N;
    const double* wb = sW + bs * N;
    const short* pa = s_ppos + a * ell;
    const short* pb = s_ppos + bs * ell;
    // one loop of (almost) the same length for every task: walk the first diagonal, then switch to the
    // second one - lanes with different diagonal lengths stay converged
    const int len0 = ell - (d0 > 0 ? d0 : -d0), len1 = d1 <= -ell ? 0 : ell - (d1 > 0 ? d1 : -d1);
    int j = d0 > 0 ? d0 : 0, k = d0 > 0 ? 0 : -d0;
    double f = d0 >= 0 ? sT[(d0 * nsig + a) * ldT + bs] : sT[(-d0 * nsig + bs) * ldT + a];
#pragma unroll 1
    for (int s2 = 0; s2 < len0 + len1; ++s2) {
      if (s2 == len0) {  // second diagonal
        j = d1 > 0 ? d1 : 0;
        k = d1 > 0 ? 0 : -d1;
        f = d1 >= 0 ? sT[(d1 * nsig + a) * ldT + bs] : sT[(-d1 * nsig + bs) * ldT + a];
      }
      put_sym(n1p, nb1, GA, GS, pa[j], pb[k], f);
      if (j + 1 < ell && k + 1 < ell) f += wa[j + C] * wb[k + C] - wa[j] * wb[k];
      ++j;
      ++k;
    }
  }
}

// ---- the solver ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
ddmpc_solve_kernel(const DdqcMpcConfig cfg, double* __restrict__ ws, int* __restrict__ work_counter,
                   const double* __restrict__ u_win, const double* __restrict__ y_win,
                   const double* __restrict__ y_r, const int32_t* __restrict__ have_prev,
                   const double* __restrict__ lamb, int8_t* __restrict__ act_state, double* __restrict__ u_opt, double* __restrict__ us_opt,
                   double* __restrict__ ys_opt, int32_t* __restrict__ status, int32_t* __restrict__ iters,
                   const int batch, long long* __restrict__ phase_clk, double* __restrict__ gram_cache, const DdqcMpcAux aux) {
  extern __shared__ __align__(16) double smem[];
  const Dims d = make_dims(cfg);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int m = d.m, p = d.p, ell = d.ell, C = d.C, N = d.N, n = d.n, L = d.L;
  const int nsig = m + p;

  double* sM = smem;                       // matrix region (aliased by window / panel / QP data)
  double* sx = smem + region_doubles(d);   // extras
  double* s_yv = sx;                       // 8*nbt : back-substitution vector
  double* s_e = s_yv + 8 * d.nbt;          // n2p   : e = D_s z on R2
  double* s_vec = s_e + d.n2p;             // scratch: inputs / result of the equilibrium QP
  double* s_red = s_vec + 2 * (d.nS + 22);  // 64
  double* s_coef = s_red + 64;              // 8: rho = sum_a coef[a] * (aug group 0 row a)
  __shared__ int s_flag;
  __shared__ int s_next;
  __shared__ int s_gmode;  // Gram of this solve: 0 regenerate, 1 cached + rank-2 correction, 2 cached as is
  __shared__ int s_state[kMaxNx];
  __shared__ int s_idx[kMaxNx + 1];
  __shared__ int s_verdict[kMaxNx];
  __shared__ short s_ppos[kMaxPos];  // Hankel row (signal a, block row j) -> position in the elimination order

  double* slot = ws + (size_t)blockIdx.x * slot_doubles(d);
  double* GA = slot;
  double* GS = slot + d.ga;

  for (int e = tid; e < nsig * ell; e += kThreads) s_ppos[e] = (short)pos_of(d, e / ell, e % ell);

  long long t_prev = clock64();
#define DDQC_PHASE(k)                                                                      \
  do {                                                                                     \
    if (phase_clk && tid == 0) {                                                           \
      const long long t_now = clock64();                                                   \
      atomicAdd((unsigned long long*)&phase_clk[k], (unsigned long long)(t_now - t_prev)); \
      t_prev = t_now;                                                                      \
    }                                                                                      \
  } while (0)
  // Gram-cache geometry (see the cached path below): panel blocks nga, non-aug trailing blocks ngs, two shared-memory
  // stages of gCH blocks behind the panel, the correction vectors a, b at the tail of the matrix region
  const int g_ntot = d.n1p + d.n2p + d.n3p + 16;
  const int gr0 = d.nb1 + d.nb2 + d.nb3, gtb0 = d.nb2 + d.nb3;
  const int nga = gr0 * d.nb1, ngs = tri(gtb0);
  const int g_cap = region_doubles(d) - 2 * g_ntot - d.ga;
  const int gCH = g_cap > 0 ? ((g_cap / 128) & ~(kWarps - 1)) : 0;
  const int g_nch = gCH > 0 ? (ngs + gCH - 1) / gCH : 0;
  const int g_off = g * 8 + ((2 * t) ^ ((g & 2) << 1));  // this lane's double2 of an 8x8 block = its DMMA accumulator fragment
  double* s_a = sM + region_doubles(d) - 2 * g_ntot;
  double* s_b = s_a + g_ntot;
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const int nb = atomicAdd(work_counter, 1);
      s_next = nb;
      s_flag = 0;
      int gm = 0;
      if (gram_cache && gCH >= kWarps && nb < batch) {
        const long long* h = reinterpret_cast<const long long*>(gram_cache + (size_t)nb * cache_stride(d));
        const int refresh = cfg.gram_refresh > 0 ? cfg.gram_refresh : kGramRefreshDefault;
        if (h[0] == 1) gm = 2;
        else if (h[0] == 2 && h[1] < refresh) gm = 1;
      }
      s_gmode = gm;
    }
    __syncthreads();
    const int b = s_next;
    if (b >= batch) break;
    if (have_prev[b] < 0) continue;  // skipped controller (an aborted evaluation run): nothing of it is touched
    const int gmode = s_gmode;
    double* gc = gram_cache ? gram_cache + (size_t)b * cache_stride(d) : nullptr;
    t_prev = clock64();
    const bool use_sr = cfg.alpha_reg_type == 0;
    // regularisation weights: per controller when `lamb` is given (grid search over lambda), else from the config
    const double lam_a = lamb ? lamb[4 * (size_t)b] : cfg.lamb_alpha, lam_s = lamb ? lamb[4 * (size_t)b + 1] : cfg.lamb_sigma;
    const double lam_as = lamb ? lamb[4 * (size_t)b + 2] : cfg.lamb_alpha_s, lam_ss = lamb ? lamb[4 * (size_t)b + 3] : cfg.lamb_sigma_s;
    const double* ub = u_win + (size_t)b * N * m;
    const double* yb = y_win + (size_t)b * N * p;
    // initial condition (the last n samples): of the data windows themselves, or of the rolling measurement arrays when
    // the Hankel data is a frozen snapshot (update_cost_threshold)
    const double* ubp = aux.u_ic ? aux.u_ic + (size_t)b * N * m : ub;
    const double* ybp = aux.y_ic ? aux.y_ic + (size_t)b * N * p : yb;
    // alpha_reg_type PREVIOUS: v_sr = H alpha_sr arrives from ddqc_ddmpc_hankel_apply; the tau row becomes tau - v_sr
    const double* vsr = (cfg.alpha_reg_type == 1 && aux.v_sr) ? aux.v_sr + (size_t)b * (nsig * ell + 1) : nullptr;

    // ---- 0. stage the (u, y) window signal-major; Gram -> workspace slot -------------------------
    DDQC_SUB(-1);
    double* sW = sM;
    double* sT = sM + nsig * N;  // first block column of G, (ell*nsig) x ldT
    // Cached G (gmode != 0).  The panel blocks of the cache are copied straight to where phase A factors the panel and
    // corrected in place; the trailing blocks stream through two stages behind the panel INSIDE phase A's trailing
    // update, which consumes them as DMMA accumulator fragments: cache -> (rank-2 correction -> cache) -> S1 -> slot,
    // so G itself never travels through the workspace slot.  Warp w owns blocks w, w + 16, ... and lane l the l-th
    // 16-byte piece of a block, in the copies and in the arithmetic alike: a thread only ever reads pieces it copied
    // itself and the pipeline needs no barrier.  The cache is read once and written once per solve: evict-first in
    // L2, so that it does not push the workspace slots, the spill area and the kernel's own instructions out of L2
    // (without the hints every later phase of the solve slows down by 5-10 %).  One SM pulls only ~25 GB/s from HBM
    // (its outstanding-request budget over the DRAM latency), ~30 k cycles for a cached G: the first two chunks are
    // requested before the correction vectors are built and arrive under the panel factorisation.  Measured and
    // dropped: a bulk L2 prefetch of the next controller's cache a third of a solve ahead - the phases it overlaps
    // (sweep, QP assembly, pivoting) lose more to the extra L2 traffic than the fetch gains.
    uint64_t pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    auto gram_fetch_panel = [&]() {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sM);
      const double* cG = gc + kCacheHdr;
      for (int bi = warp; bi < nga; bi += kWarps)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + 8u * (uint32_t)(bi * 64 + g_off)), "l"(cG + (size_t)bi * 64 + g_off), "l"(pol_stream) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto gram_fetch = [&](int c) {  // trailing chunk c -> stage c & 1 (an empty group when c is past the end)
      const int c0 = c * gCH, c1 = c0 + gCH < ngs ? c0 + gCH : ngs;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sM + d.ga + (c & 1) * gCH * 64);
      const double* cG = gc + kCacheHdr + d.ga;
      for (int bb = c0 + warp; bb < c1; bb += kWarps)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + 8u * (uint32_t)((bb - c0) * 64 + g_off)), "l"(cG + (size_t)bb * 64 + g_off), "l"(pol_stream) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (gmode == 0) {
      for (int e = tid; e < N * m; e += kThreads) sW[(e % m) * N + e / m] = ub[e];
      for (int e = tid; e < N * p; e += kThreads) sW[(m + e % p) * N + e / p] = yb[e];
    } else {
      // a (old first Hankel column) and b (new last column) by position in the elimination order; pad and aug
      // positions stay 0, the ones row has a = b = 1 (its entries move by b_P - a_P)
      gram_fetch_panel();
      gram_fetch(0);
      gram_fetch(1);
      for (int e = tid; e < 2 * g_ntot; e += kThreads) s_a[e] = 0.0;
      __syncthreads();
      if (gmode == 1) {
        for (int e = tid; e < nsig * ell; e += kThreads) {
          const int a = e / ell, j = e % ell, P = s_ppos[e];
          const double* w = a < m ? ub + a : yb + (a - m);
          const int st = a < m ? m : p;
          s_a[P] = j == 0 ? gc[2 + a] : w[(size_t)(j - 1) * st];
          s_b[P] = w[(size_t)(j + C - 1) * st];
        }
        if (tid == 0) { s_a[n * m] = 1.0; s_b[n * m] = 1.0; }
      }
    }
    // aug block rows and pad rows: explicit zeros first
    {
      const int r0 = d.nb1 + d.nb2 + d.nb3;  // first aug block row (all-rows numbering)
      double2* za = reinterpret_cast<double2*>((gmode ? sM : GA) + (size_t)r0 * d.nb1 * 64);
      for (int e = tid; e < 2 * d.nb1 * 32; e += kThreads) za[e] = make_double2(0.0, 0.0);
      const int tb = d.nb2 + d.nb3;
      double2* zs = reinterpret_cast<double2*>(GS + (size_t)tri(tb) * 64);
      const int cnt = (tri(tb + 2) - tri(tb)) * 32;
      for (int e = tid; e < cnt; e += kThreads) zs[e] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    DDQC_SUB(0);
    if (gmode != 0) {
      // panel blocks of the cached G: rank-2 correction in place (shared memory) and back to the cache
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      if (gmode == 1) {
        double* cG = gc + kCacheHdr;
        int row = warp / d.nb1, col = warp % d.nb1;
        for (int bi = warp; bi < nga; bi += kWarps) {
          const int hi = row * 8 + g, lo = col * 8 + 2 * t;
          double2* pv = reinterpret_cast<double2*>(sM + (size_t)bi * 64 + g_off);
          double2 v = *pv;
          const double ah = s_a[hi], bh = s_b[hi];
          const double2 al = *reinterpret_cast<const double2*>(s_a + lo), bl = *reinterpret_cast<const double2*>(s_b + lo);
          v.x = fma(bh, bl.x, fma(-ah, al.x, v.x));
          v.y = fma(bh, bl.y, fma(-ah, al.y, v.y));
          asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(cG + (size_t)bi * 64 + g_off), "d"(v.x), "d"(v.y), "l"(pol_stream) : "memory");
          *pv = v;
          col += kWarps;
          while (col >= d.nb1) { col -= d.nb1; ++row; }
        }
      }
    } else {
      // pad rows (identity) of the three classes
      const int npad = (d.n1p - d.n1) + (d.n2p - d.n2) + (d.n3p - d.n3);
      const int ntot = d.n1p + d.n2p + d.n3p + 16;
      for (int e = tid; e < npad * ntot; e += kThreads) {
        int k = e / ntot;
        const int q = e % ntot;
        int P;
        if (k < d.n1p - d.n1) P = d.n1 + k;
        else if ((k -= d.n1p - d.n1) < d.n2p - d.n2) P = d.n1p + d.n2 + k;
        else P = d.n1p + d.n2p + d.n3 + (k - (d.n2p - d.n2));
        put_sym(d, GA, GS, P, q, P == q ? 1.0 : 0.0);
      }
      DDQC_SUB(1);
      // First block column of G on the tensor cores:  T[(j,a)][b] = sum_c w_a[j+c] w_b[c]  (b = nsig
      // is the all-ones signal), i.e. [Hankel rows] x [first block row]^T, K = C.  Operands are read
      // straight from the window, H is never materialised.
      {
        const int nrho = ell * nsig, nrb = (nrho + 7) >> 3, ncb = (nsig + 1 + 7) >> 3, ldT = ncb * 8;
        // a warp takes two row blocks at a time against the same B fragments (four independent accumulator chains)
        for (int cb = 0; cb < ncb; ++cb) {
          const int bsig = cb * 8 + g;
          const bool bv = bsig < nsig, bone = bsig == nsig;
          const double* wb = sW + (bv ? bsig * N : 0);
          for (int rb0 = warp; rb0 < nrb; rb0 += 2 * kWarps) {
            const int rho0 = rb0 * 8 + g, rho1 = (rb0 + kWarps) * 8 + g;
            const bool av0 = rho0 < nrho, av1 = rho1 < nrho;
            const double* wa0 = sW + (av0 ? (rho0 % nsig) * N + rho0 / nsig : 0);
            const double* wa1 = sW + (av1 ? (rho1 % nsig) * N + rho1 / nsig : 0);
            double2 p0 = {0.0, 0.0}, p1 = {0.0, 0.0}, q0 = {0.0, 0.0}, q1 = {0.0, 0.0};
#pragma unroll 2
            for (int c0 = 0; c0 < C; c0 += 8) {
              const int ca = c0 + t, cbb = c0 + 4 + t;
              const bool va = ca < C, vb = cbb < C;
              const double b0 = va ? (bv ? wb[ca] : (bone ? 1.0 : 0.0)) : 0.0;
              const double b1 = vb ? (bv ? wb[cbb] : (bone ? 1.0 : 0.0)) : 0.0;
              dmma(p0, (av0 && va) ? wa0[ca] : 0.0, b0);
              dmma(p1, (av0 && vb) ? wa0[cbb] : 0.0, b1);
              dmma(q0, (av1 && va) ? wa1[ca] : 0.0, b0);
              dmma(q1, (av1 && vb) ? wa1[cbb] : 0.0, b1);
            }
            if (av0) {
              sT[rho0 * ldT + cb * 8 + 2 * t] = p0.x + p1.x;
              sT[rho0 * ldT + cb * 8 + 2 * t + 1] = p0.y + p1.y;
            }
            if (av1) {
              sT[rho1 * ldT + cb * 8 + 2 * t] = q0.x + q1.x;
              sT[rho1 * ldT + cb * 8 + 2 * t + 1] = q0.y + q1.y;
            }
          }
        }
        __syncthreads();
        DDQC_SUB(2);
        gram_walk(d.n1p, d.nb1, N, ell, C, nsig, sW, sT, ldT, s_ppos, GA, GS);
        DDQC_SUB(3);
        // ones row: G[(j,a), one] = sum_c w_a[j+c] = T[(j,a)][nsig];  G[one][one] = C
        const int Pone = n * m;
        for (int e = tid; e < nrho; e += kThreads) put_sym(d, GA, GS, s_ppos[(e % nsig) * ell + e / nsig], Pone, sT[e * ldT + nsig]);
        if (tid == 0) put_sym(d, GA, GS, Pone, Pone, (double)C);
      }
    }
    {
      const int Pone = n * m;
      double* GA = gmode ? sM : slot;  // cached path: the panel already lives in shared memory
      // aug rows.  group 0 (block row tb): A_u (m: ones on ALL input rows of a component), A_y (p), e_one, tau;
      //            group 1 (block row tb + 1): T_u (m: ones on the terminal input rows), T_y (p: free + terminal outputs)
      const int Pa0 = d.n1p + d.n2p + d.n3p, Pa1 = Pa0 + 8;
      for (int e = tid; e < nsig * ell; e += kThreads) {
        const int a = e / ell, j = e % ell;
        const int P = pos_of(d, a, j);
        put_sym(d, GA, GS, Pa0 + a, P, 1.0);
        double tv = 0.0;
        if (a < m) {
          if (j >= L) put_sym(d, GA, GS, Pa1 + a, P, 1.0);
          if (j < n) tv = ubp[(size_t)(N - n + j) * m + a];
        } else {
          if (j >= n) put_sym(d, GA, GS, Pa1 + a, P, 1.0);
          if (j < n) tv = ybp[(size_t)(N - n + j) * p + (a - m)];
        }
        if (vsr) tv -= vsr[e];
        if (j < n || vsr) put_sym(d, GA, GS, Pa0 + nsig + 1, P, tv);
      }
      if (tid == 0) {
        put_sym(d, GA, GS, Pa0 + nsig, Pone, 1.0);
        put_sym(d, GA, GS, Pa0 + nsig + 1, Pone, vsr ? 1.0 - vsr[nsig * ell] : 1.0);
      }
    }
    __syncthreads();
    if (gc) {
      if (gmode == 0) {  // regenerated: G (slot layout) -> cache
        const double2* src = reinterpret_cast<const double2*>(slot);
        double2* dst = reinterpret_cast<double2*>(gc + kCacheHdr);
        const int n2 = (int)(slot_doubles(d) / 2);
#pragma unroll 4
        for (int e = tid; e < n2; e += kThreads) __stcs(dst + e, __ldcg(src + e));
      }
      if (tid == 0) {
        long long* h = reinterpret_cast<long long*>(gc);
        h[1] = gmode == 0 ? 0 : h[1] + (gmode == 1 ? 1 : 0);
        h[0] = 1;
      }
    }
    DDQC_SUB(4);
    DDQC_PHASE(0);

    // ---- A. eliminate R1: panel in shared memory, S1 = trailing - L L^T through global ------------
    if (gmode == 0) copy_slot_to_smem(sM, GA, d.ga / 2);
    __syncthreads();
    factor_columns<true>(sM, d.nb1, d.nbA, d.nb1, &s_flag);
    DDQC_PHASE(1);
    if (gmode != 0) {
      // cached path: trailing blocks cache -> stage -> (correction -> cache) -> S1 = G - L L' -> slot
      double* cG = gc + kCacheHdr + d.ga;
      int bb = warp, I = 0, J = 0;
      {
        I = (int)((sqrtf(8.0f * bb + 1.0f) - 1.0f) * 0.5f);
        while (tri(I + 1) <= bb) ++I;
        while (tri(I) > bb) --I;
        J = bb - tri(I);
      }
      for (int c = 0; c < g_nch; ++c) {
        if (c + 1 < g_nch) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        const int c0 = c * gCH, c1 = c0 + gCH < ngs ? c0 + gCH : ngs;
        const double* stage = sM + d.ga + (c & 1) * gCH * 64;
        for (; bb < c1; bb += kWarps) {
          double2 v = *reinterpret_cast<const double2*>(stage + (bb - c0) * 64 + g_off);
          if (gmode == 1) {
            const int hi = d.n1p + I * 8 + g, lo = d.n1p + J * 8 + 2 * t;
            const double ah = s_a[hi], bh = s_b[hi];
            const double2 al = *reinterpret_cast<const double2*>(s_a + lo), bl = *reinterpret_cast<const double2*>(s_b + lo);
            v.x = fma(bh, bl.x, fma(-ah, al.x, v.x));
            v.y = fma(bh, bl.y, fma(-ah, al.y, v.y));
            asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(cG + (size_t)bb * 64 + g_off), "d"(v.x), "d"(v.y), "l"(pol_stream) : "memory");
          }
          const double* pi = sM + (size_t)(d.nb1 + I) * d.nb1 * 64;
          const double* pj = sM + (size_t)(d.nb1 + J) * d.nb1 * 64;
          double2 a0 = {0.0, 0.0}, a1 = {0.0, 0.0};
          for (int Jc = 0; Jc < d.nb1; ++Jc) {
            dmma(a0, ld_ab(pi + Jc * 64, g, t, 0), ld_ab(pj + Jc * 64, g, t, 0));
            dmma(a1, ld_ab(pi + Jc * 64, g, t, 1), ld_ab(pj + Jc * 64, g, t, 1));
          }
          v.x -= a0.x + a1.x;
          v.y -= a0.y + a1.y;
          *reinterpret_cast<double2*>(GS + (size_t)bb * 64 + g_off) = v;
          J += kWarps;
          while (J > I) { J -= I + 1; ++I; }
        }
        gram_fetch(c + 2);  // into the stage just consumed
      }
    }
    {
      const int nblk = tri(d.nbt + 1);
      for (int bb = (gmode ? ngs : 0) + warp; bb < nblk; bb += 2 * kWarps) {
        const int b2 = bb + kWarps;
        const bool two = b2 < nblk;
        int i0 = (int)((sqrtf(8.0f * bb + 1.0f) - 1.0f) * 0.5f);
        while (tri(i0 + 1) <= bb) ++i0;
        while (tri(i0) > bb) --i0;
        const int j0 = bb - tri(i0);
        int i1 = i0, j1 = j0;
        if (two) {
          i1 = (int)((sqrtf(8.0f * b2 + 1.0f) - 1.0f) * 0.5f);
          while (tri(i1 + 1) <= b2) ++i1;
          while (tri(i1) > b2) --i1;
          j1 = b2 - tri(i1);
        }
        double2 c0 = ld_c(GS + (size_t)bb * 64, g, t);
        double2 c1 = two ? ld_c(GS + (size_t)b2 * 64, g, t) : make_double2(0.0, 0.0);
        double2 a0 = {0.0, 0.0}, a1 = {0.0, 0.0};
        const double* pi0 = sM + (size_t)(d.nb1 + i0) * d.nb1 * 64;
        const double* pj0 = sM + (size_t)(d.nb1 + j0) * d.nb1 * 64;
        const double* pi1 = sM + (size_t)(d.nb1 + i1) * d.nb1 * 64;
        const double* pj1 = sM + (size_t)(d.nb1 + j1) * d.nb1 * 64;
        for (int J = 0; J < d.nb1; ++J) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            dmma(a0, ld_ab(pi0 + J * 64, g, t, h), ld_ab(pj0 + J * 64, g, t, h));
            if (two) dmma(a1, ld_ab(pi1 + J * 64, g, t, h), ld_ab(pj1 + J * 64, g, t, h));
          }
        }
        c0.x -= a0.x; c0.y -= a0.y;
        st_c(GS + (size_t)bb * 64, g, t, c0);
        if (two) {
          c1.x -= a1.x; c1.y -= a1.y;
          st_c(GS + (size_t)b2 * 64, g, t, c1);
        }
      }
    }
    __syncthreads();
    DDQC_PHASE(2);

    // ---- B. alpha_sr = alpha of the optimal reachable equilibrium for y_r -----------------------------
    //   (u_sr, y_sr) = argmin lamb_alpha_s rho' (G + D_s)^-1 rho + |y_s - y_r|_S^2, u_s in U_s,
    //   rho = A_u u_s + A_y y_s + e_one;   v_sr = H alpha_sr = rho - e,  e = D_s (G + D_s)^-1 rho on R2
    const int ncol_all = d.nb2 + d.nb3, tb = ncol_all;        // tb: the aug block row held in shared memory
    const int aug_t = nsig;                                   // phase C: aug index of the c-row (T_u, T_y before it)
    const int Paug = d.n2p + d.n3p;                           // trailing position of the first aug row
    if (tid < 8) s_coef[tid] = 0.0;
    for (int q = tid; q < d.n2p; q += kThreads) s_e[q] = 0.0;
    __syncthreads();
    if (use_sr) {
      const double ds = lam_as / lam_ss;
      copy_slot_to_smem(sM, GS, d.gs / 2);  // block rows 0..tb: matrix + aug group 0
      __syncthreads();
      for (int q = tid; q < d.n2; q += kThreads) *elem_tri(sM, q, q) += ds;
      __syncthreads();
      factor_columns<false>(sM, 0, d.nbt, ncol_all, &s_flag);
      schur_trailing(sM, d.nbt, ncol_all);  // aug x aug block: -K_s' (G + D_s)^-1 K_s
      __syncthreads();
      DDQC_PHASE(3);
      DDQC_SUB(-1);
      if (cfg.alpha_sr_anchor == 1) {
        // alpha_sr anchored at the PREVIOUS solve's optimal equilibrium (the other reading of the package's
        // "approximation of alpha_Lin^sr(D_t)"): rho = A_u u_s^prev + A_y y_s^prev + e_one with the values the previous
        // launch left in us_opt / ys_opt; a controller's first solve has no previous equilibrium and uses alpha_sr = 0
        if (tid <= nsig) {
          double v = 0.0;
          if (have_prev[b] > 0) v = tid < m ? us_opt[(size_t)b * m + tid] : (tid < nsig ? ys_opt[(size_t)b * p + (tid - m)] : 1.0);
          s_coef[tid] = v;
        }
      } else {
        // 6-variable box QP for the equilibrium (bounds on u_s only): H 36 | f 6 | lo 6 | hi 6 | x 6 at s_vec (free
        // at this point), assembled by 42 threads, every active set tested by its own thread
        double* io = s_vec;
        if (tid < 36) {
          const int a = tid / 6, c = tid % 6;
          double v = (a < nsig && c < nsig) ? -2.0 * lam_as * *elem_tri(sM, Paug + (a > c ? a : c), Paug + (a > c ? c : a)) : 0.0;
          if (a == c && a >= m && a < nsig) v += 2.0 * cfg.S[a - m];
          io[tid] = v;
        } else if (tid < 42) {
          const int a = tid - 36;
          double v = a < nsig ? -2.0 * lam_as * *elem_tri(sM, Paug + nsig, Paug + a) : 0.0;
          if (a >= m && a < nsig) v -= 2.0 * cfg.S[a - m] * y_r[(size_t)b * p + (a - m)];
          io[36 + a] = v;
          io[42 + a] = a < m ? cfg.Us_lo[a] : -INFINITY;
          io[48 + a] = a < m ? cfg.Us_hi[a] : INFINITY;
        }
        if (tid == 0) s_next = 0x7fffffff;  // the work index was copied to `b`; reused as the winner slot
        __syncthreads();
        int ncombo = 1;
        for (int i = 0; i < m; ++i) ncombo *= 3;
        double xq[6];
        if (tid < ncombo && ncombo <= kThreads) {
          if (small_box_qp_combo(io, nsig, m, tid, xq)) atomicMin(&s_next, tid);
        }
        __syncthreads();
        const int win = s_next;
        if (win == 0x7fffffff) {  // no set passed the tolerances (or m too large): serial pivoting
          if (tid == 0) {
            if (!small_box_qp(io, nsig)) s_flag = 1;
            for (int a = 0; a < nsig; ++a) s_coef[a] = io[54 + a];
            s_coef[nsig] = 1.0;
          }
        } else if (tid == win) {
          for (int a = 0; a < 6; ++a)
            if (a < nsig) s_coef[a] = xq[a];
          s_coef[nsig] = 1.0;
        }
      }
      __syncthreads();
      DDQC_SUB(7);
      // y = L^-1 rho = combination of the forward-substituted aug rows; z = L^-T y
      for (int q = tid; q < 8 * ncol_all; q += kThreads) {
        double acc = 0.0;
        for (int a = 0; a <= nsig; ++a) acc += s_coef[a] * *elem_tri(sM, Paug + a, q);
        s_yv[q] = acc;
      }
      __syncthreads();
      backsub_blocked(sM, ncol_all, s_yv);
      for (int q = tid; q < d.n2p; q += kThreads) s_e[q] = q < d.n2 ? ds * s_yv[q] : 0.0;
      __syncthreads();
      if (aux.sr_out) {  // (u_sr, y_sr, e) for ddqc_ddmpc_recover: v_sr = rho - e
        double* so = aux.sr_out + (size_t)b * (nsig + d.n2);
        for (int q = tid; q < nsig + d.n2; q += kThreads) so[q] = q < nsig ? s_coef[q] : s_e[q - nsig];
      }
      DDQC_PHASE(4);
    }

    // ---- C. main system: S1 + D_0, aug rows (T_u, T_y, c), c = tau - rho + e; eliminate R2 ------------
    {
      // block rows 0..tb-1 (asynchronous: the aug block row below is assembled while they arrive)
      {
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(sM);
        const double2* s2 = reinterpret_cast<const double2*>(GS);
        for (int e = tid; e < tri(tb) * 32; e += kThreads)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + 16u * (uint32_t)e), "l"(s2 + e) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      // aug block row: T rows from workspace block row tb+1, the c-row combined from block row tb
      const double* g0 = GS + (size_t)tri(tb) * 64;       // aug group 0 blocks (tb, J)
      const double* g1 = GS + (size_t)tri(tb + 1) * 64;   // aug group 1 blocks (tb + 1, J)
      double* sa = sM + (size_t)tri(tb) * 64;
      for (int e = tid; e < tb * 64; e += kThreads) {
        const int J = e >> 6, a = (e >> 3) & 7, c = e & 7, q = 8 * J + c;
        double v = 0.0;
        if (a < nsig) v = g1[J * 64 + eoff(a, c)];
        else if (a == nsig) {
          v = g0[J * 64 + eoff(nsig + 1, c)];
          for (int k = 0; k <= nsig; ++k) v -= s_coef[k] * g0[J * 64 + eoff(k, c)];
          if (q < d.n2) v += s_e[q];
        }
        sa[J * 64 + eoff(a, c)] = v;
      }
      if (tid < 64) {
        const int a = tid >> 3, c = tid & 7;
        const double* b00 = g0 + (size_t)tb * 64;        // (tb, tb)     : group 0 x group 0 (lower)
        const double* b10 = g1 + (size_t)tb * 64;        // (tb+1, tb)   : group 1 rows x group 0 columns
        const double* b11 = g1 + (size_t)(tb + 1) * 64;  // (tb+1, tb+1) : group 1 x group 1 (lower)
        const int hi = a > c ? a : c, lo = a > c ? c : a;
        double v = 0.0;
        if (hi < nsig) v = b11[eoff(hi, lo)];
        else if (hi == nsig && lo < nsig) {  // c-row against a T row
          v = b10[eoff(lo, nsig + 1)];
          for (int k = 0; k <= nsig; ++k) v -= s_coef[k] * b10[eoff(lo, k)];
        } else if (hi == nsig && lo == nsig) {  // w' B w, w = (-coef[0..nsig], +1 on tau)
          for (int i = 0; i <= nsig + 1; ++i) {
            const double wi = i <= nsig ? -s_coef[i] : 1.0;
            for (int k = 0; k <= nsig + 1; ++k) {
              const double wk = k <= nsig ? -s_coef[k] : 1.0;
              v += wi * wk * b00[eoff(i > k ? i : k, i > k ? k : i)];
            }
          }
        }
        sa[tb * 64 + eoff(a, c)] = v;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    for (int q = tid; q < d.n2; q += kThreads) {
      const int j = q / p, comp = q % p;
      const double w = (j < n || j >= L) ? lam_s : cfg.Q[comp] * lam_s / (cfg.Q[comp] + lam_s);
      *elem_tri(sM, q, q) += lam_a / w;
    }
    __syncthreads();
    factor_columns<false>(sM, 0, d.nbt, d.nb2, &s_flag);
    schur_trailing(sM, d.nbt, d.nb2);
    __syncthreads();
    DDQC_PHASE(5);

    // ---- sweep: S7 (dense, lower) <- Schur complement on [R3 (n3) | T_u, T_y, c] ---------------------
    // ---- Z = K'(G + D_0)^-1 K from the Schur complement S on [R3 | aug] WITHOUT a scalar Gauss-Jordan sweep:
    // append an identity block E below S, [R3 (nb3) | aug (1) | E (nb3)], eliminate R3 once more on the tensor
    // cores; the trailing blocks are then  (aug,aug) = S_aa - S_aB S_BB^-1 S_Ba = -Z_aa,  (E,aug) = -S_BB^-1 S_Ba
    // = -Z_Ba,  (E,E) = -S_BB^-1 = -Z_BB:   Z = -trailing, one sign for all blocks.
    const int nr2 = 2 * d.nb3 + 1;
    repack_schur(sM, d.nb2, d.nb3);
    __syncthreads();
    factor_columns<false>(sM, 0, nr2, d.nb3, &s_flag);
    schur_trailing(sM, nr2, d.nb3);
    __syncthreads();
    DDQC_PHASE(6);

    // ---- box QP assembly: Hq = 2 lamb_alpha Z + structure, f ----------------------------------------
    const int nx = d.nx, ldq = d.ldq;
    double* Hq = sM + qp_scratch_doubles(d);
    double* qv = Hq + nx * ldq;  // 20 vectors of kMaxNx
    QpWork q;
    q.Hq = Hq; q.ldq = ldq; q.nx = nx;
    q.F = sM; q.ldf = 0;
    q.f = qv; q.lo = qv + kMaxNx; q.hi = qv + 2 * kMaxNx; q.x = qv + 3 * kMaxNx; q.g = qv + 4 * kMaxNx;
    q.rhs = qv + 5 * kMaxNx; q.rs = qv + 6 * kMaxNx; q.sl = qv + 7 * kMaxNx; q.su = qv + 8 * kMaxNx;
    q.zl = qv + 9 * kMaxNx; q.zu = qv + 10 * kMaxNx; q.dx = qv + 11 * kMaxNx; q.dzl = qv + 12 * kMaxNx;
    q.dzu = qv + 13 * kMaxNx; q.rd = qv + 14 * kMaxNx; q.dg = qv + 15 * kMaxNx;
    q.s_red = s_red; q.state = s_state; q.idx = s_idx; q.verdict = s_verdict; q.s_flag = &s_flag;
    DDQC_SUB(-1);
    assemble_hq(sM, Hq, q.f, -2.0 * lam_a, d.nb3, d.n3, nx, ldq);  // Z = -trailing
    __syncthreads();
    DDQC_SUB(5);
    {
      // structure of the tracking cost: R on (u_k - u_s) for the L-n predicted inputs and, with u_k = u_s, on the n
      // terminal inputs against nothing / on (u_s - u_past) ... in short every |.|_R, |.|_Q, |.|_S term that is not a
      // function of alpha alone.  One thread per predicted input, one warp per equilibrium component (its sum over
      // the n past samples is a strided warp reduction instead of n dependent global loads by one thread).
      const int nB = d.n3;
      if (tid < nB) {
        const int comp = tid % m;
        const double w2 = 2.0 * cfg.R[comp];
        Hq[tid * ldq + tid] += w2;
        Hq[tid * ldq + nB + comp] -= w2;
        Hq[(nB + comp) * ldq + tid] -= w2;
      }
      if (warp < nsig) {
        const int i = warp;
        double sum = 0.0;
        for (int j = lane; j < n; j += 32)
          sum += i < m ? ubp[(size_t)(N - n + j) * m + i] : ybp[(size_t)(N - n + j) * p + (i - m)];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
          const int qq = nB + i;
          if (i < m) {
            Hq[qq * ldq + qq] += 2.0 * (L - n) * cfg.R[i] + 2.0 * n * cfg.R[i];
            q.f[qq] += -2.0 * cfg.R[i] * sum;
          } else {
            const int c = i - m;
            Hq[qq * ldq + qq] += 2.0 * (n * cfg.Q[c] + cfg.S[c]);
            q.f[qq] += -2.0 * (cfg.Q[c] * sum + cfg.S[c] * y_r[(size_t)b * p + c]);
          }
        }
      }
    }
    const bool warm = act_state != nullptr && have_prev[b] != 0;
    if (tid < nx) {
      double lo, hi;
      if (tid < d.n3) {
        lo = cfg.U_lo[tid % m];
        hi = cfg.U_hi[tid % m];
      } else if (tid < d.n3 + m) {
        const int comp = tid - d.n3;
        lo = fmax(cfg.U_lo[comp], cfg.Us_lo[comp]);
        hi = fmin(cfg.U_hi[comp], cfg.Us_hi[comp]);
      } else {
        lo = -INFINITY;
        hi = INFINITY;
      }
      q.lo[tid] = lo;
      q.hi[tid] = hi;
      int st = 0;
      if (warm) {
        // receding horizon: predicted input k of this solve was input k+1 of the previous one
        const int8_t* as = act_state + (size_t)b * nx;
        if (tid < d.n3) st = tid + m < d.n3 ? as[tid + m] : as[tid];
        else st = as[tid];
        if ((st < 0 && !isfinite(lo)) || (st > 0 && !isfinite(hi))) st = 0;
      }
      s_state[tid] = st;
    }
    __syncthreads();
    DDQC_SUB(6);
    DDQC_PHASE(7);

    // ---- D. box QP ------------------------------------------------------------------------------------
    int passes = 0, ipm_it = 0, converged;
    {
      const double fmax_abs = block_reduce(tid < nx ? fabs(q.f[tid]) : 0.0, RED_MAX, s_red);
      const double tol_g = 1e-11 * fmax(1.0, fmax_abs);
      const int cap = cfg.max_iter > 0 ? cfg.max_iter : 12;
      converged = bpp_solve(q, tol_g, warm ? cap : (cap < 8 ? cap : 8), &passes);
      DDQC_PHASE(8);
      if (!converged) {
        ipm_it = ipm_solve(q, 60);
        // keep the interior-point iterate in case the polish does not settle
        if (tid < nx) q.dx[tid] = q.x[tid];
        __syncthreads();
        converged = bpp_solve(q, tol_g, 6, &passes);
        if (!converged) {
          if (tid < nx) q.x[tid] = q.dx[tid];
          converged = ipm_it > 0;
          __syncthreads();
        }
        ipm_it = (ipm_it < 0 ? -ipm_it : ipm_it) - 1;
        DDQC_PHASE(9);
      }
    }

    // ---- outputs ----------------------------------------------------------------------------------------
    for (int e = tid; e < (L + 1) * m; e += kThreads) {
      const int k = e / m, comp = e % m;
      u_opt[(size_t)b * (L + 1) * m + e] = (k < L - n) ? q.x[k * m + comp] : q.x[d.n3 + comp];
    }
    for (int i = tid; i < m; i += kThreads) us_opt[(size_t)b * m + i] = q.x[d.n3 + i];
    for (int i = tid; i < p; i += kThreads) ys_opt[(size_t)b * p + i] = q.x[d.n3 + m + i];
    if (act_state != nullptr && tid < nx) act_state[(size_t)b * nx + tid] = (int8_t)s_state[tid];
    if (tid == 0) {
      status[b] = s_flag ? 2 : (converged ? 0 : 1);
      iters[b] = passes + (ipm_it << 16);
    }
    DDQC_PHASE(10);
  }
}

__global__ void ddmpc_store_kernel(const DdqcMpcConfig cfg, double* __restrict__ u_win,
                                   double* __restrict__ y_win, const double* __restrict__ u_new,
                                   const double* __restrict__ y_new, const int batch, double* __restrict__ gram_cache) {
  // one CTA per controller: shift the window one sample (in place, ordered) and append
  const int N = cfg.N, m = cfg.m, p = cfg.p;
  const int b = blockIdx.x;
  if (b >= batch) return;
  double* ub = u_win + (size_t)b * N * m;
  double* yb = y_win + (size_t)b * N * p;
  if (gram_cache) cache_note_shift(gram_cache, make_dims(cfg), b, ub, yb);
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

// ---- closed-loop evaluation helpers (controller_evaluation.py:345-450 of the reference, batched) ----

// env action (normalised to [-1, 1], fp32) from the optimal input of step `n_step`:
// linear_interpolate(u, lo, hi, -1, 1) = -1 + (u - lo) * 2 / (hi - lo)   (math_utils.py:13-20 order)
__global__ void ddmpc_action_kernel(const double* __restrict__ u_opt, const int Lp1, const int m, const int n_step,
                                    const float* __restrict__ bounds /* [m][2] */, float* __restrict__ action,
                                    double* __restrict__ u_applied, const int batch) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= batch * m) return;
  const int b = e / m, i = e % m;
  const double u = u_opt[((size_t)b * Lp1 + n_step) * m + i];
  const float lo = bounds[2 * i], hi = bounds[2 * i + 1];
  const float uf = (float)u;
  action[e] = -1.0f + (uf - lo) * 2.0f / (hi - lo);
  u_applied[e] = u;
}

// store_input_output_measurement + tracking-error bookkeeping in one launch (one CTA per controller):
// shifts the windows, appends (u_applied, y), accumulates |y - y_r|^2 and raises the early-stop flag
// when |y - y_r| - d0 > max_inc (controller_evaluation.py:424-447).  Controllers that already stopped
// (fail_step >= 0) keep their windows and sums frozen, like an aborted evaluation.
__global__ void ddmpc_post_step_kernel(const DdqcMpcConfig cfg, double* __restrict__ gram_cache, double* __restrict__ u_win,
                                       double* __restrict__ y_win, const double* __restrict__ u_new,
                                       const float* __restrict__ y_meas, const long long y_rs, const long long y_cs,
                                       const double* __restrict__ y_r, const double* __restrict__ d0,
                                       const double max_inc, double* __restrict__ sq_sum, int32_t* __restrict__ n_acc,
                                       int32_t* __restrict__ fail_step, double* __restrict__ y_log, const int step,
                                       const int batch) {
  const int N = cfg.N, m = cfg.m, p = cfg.p;
  const int b = blockIdx.x;
  if (b >= batch) return;
  if (fail_step[b] >= 0) return;  // uniform per CTA
  double* ub = u_win + (size_t)b * N * m;
  double* yb = y_win + (size_t)b * N * p;
  if (gram_cache) cache_note_shift(gram_cache, make_dims(cfg), b, ub, yb);
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
  if (threadIdx.x < p) {
    const double y = (double)y_meas[(long long)b * y_rs + threadIdx.x * y_cs];
    yb[(size_t)(N - 1) * p + threadIdx.x] = y;
    if (y_log) y_log[((size_t)step * batch + b) * p + threadIdx.x] = y;
  }
  if (threadIdx.x == 0) {
    double sq = 0.0;
    for (int i = 0; i < p; ++i) {
      const double e = (double)y_meas[(long long)b * y_rs + i * y_cs] - y_r[(size_t)b * p + i];
      sq += e * e;
    }
    sq_sum[b] += sq;
    n_acc[b] += 1;
    if (max_inc >= 0.0 && sqrt(sq) - d0[b] > max_inc) fail_step[b] = step;
  }
}

size_t smem_bytes(const Dims& d) { return (size_t)(region_doubles(d) + extra_doubles(d)) * sizeof(double); }

int check_cfg(const DdqcMpcConfig* c) {
  if (!c) return DDQC_E_NULL;
  if (c->m < 1 || c->p < 1 || c->m > kMaxSig || c->p > kMaxSig || c->n < 1 || c->L <= c->n) return DDQC_E_SHAPE;
  const Dims d = make_dims(*c);
  if (d.C < 1 || d.nx > kMaxNx || (d.m + d.p) * d.ell > kMaxPos || d.m + d.p + 2 > 8) return DDQC_E_SHAPE;
  if (smem_bytes(d) + 6144 > (size_t)kSmemLimit) return DDQC_E_SHAPE;  // 6144 >= static shared memory of the kernel
  if (c->alpha_reg_type < 0 || c->alpha_reg_type > 2) return DDQC_E_CONFIG;
  if (c->alpha_sr_anchor != 0 && c->alpha_sr_anchor != 1) return DDQC_E_CONFIG;
  return 0;
}

}  // namespace

static long long* g_phase_clk = nullptr;
// Developer aid (not part of the reference-facing surface): device buffer of 12 cycle counters that the
// solver accumulates per phase.  Pass NULL to switch it off.
extern "C" int ddqc_ddmpc_debug_phase_buffer(void* dev_ptr) {
  g_phase_clk = static_cast<long long*>(dev_ptr);
#ifdef DDQC_PROF_FACTOR
  cudaMemcpyToSymbol(g_fprof, &g_phase_clk, sizeof(void*));
#endif
#ifdef DDQC_PROF_SUB
  cudaMemcpyToSymbol(g_subclk, &g_phase_clk, sizeof(void*));
#endif

  return 0;
}

extern "C" int64_t ddqc_ddmpc_workspace_bytes(const DdqcMpcConfig* cfg, int64_t batch) {
  if (check_cfg(cfg) != 0 || batch < 0) return -1;
  const Dims d = make_dims(*cfg);
  const int64_t slots = batch < kMaxSlots ? batch : kMaxSlots;
  return 256 + (int64_t)(slot_doubles(d) * sizeof(double)) * (slots > 0 ? slots : 1);
}

extern "C" int64_t ddqc_ddmpc_gram_cache_bytes(const DdqcMpcConfig* cfg, int64_t batch) {
  if (check_cfg(cfg) != 0 || batch < 0) return -1;
  return (int64_t)(cache_stride(make_dims(*cfg)) * sizeof(double)) * batch;
}

extern "C" int ddqc_ddmpc_solve(const DdqcMpcConfig* cfg, void* ws, const double* u_win, const double* y_win,
                                const double* y_r,
                                const int32_t* have_prev, const double* lamb, int8_t* act_state, double* u_opt,
                                double* us_opt,
                                double* ys_opt, int32_t* status, int32_t* iters, int64_t batch, void* gram_cache,
                                void* stream) {
  return ddqc_ddmpc_solve_ex(cfg, ws, u_win, y_win, y_r, have_prev, lamb, act_state, u_opt, us_opt, ys_opt, status, iters,
                             batch, gram_cache, nullptr, stream);
}

extern "C" int ddqc_ddmpc_solve_ex(const DdqcMpcConfig* cfg, void* ws, const double* u_win, const double* y_win,
                                   const double* y_r, const int32_t* have_prev, const double* lamb, int8_t* act_state,
                                   double* u_opt, double* us_opt, double* ys_opt, int32_t* status, int32_t* iters,
                                   int64_t batch, void* gram_cache, const DdqcMpcAux* aux_host, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  DdqcMpcAux aux = {nullptr, nullptr, nullptr, nullptr};
  if (aux_host) aux = *aux_host;
  if ((aux.u_ic == nullptr) != (aux.y_ic == nullptr)) return DDQC_E_NULL;
  // PREVIOUS needs v_sr (zeros on a first solve); a frozen snapshot has no one-sample shift for the Gram cache to follow
  if (cfg->alpha_reg_type == 1 && !aux.v_sr) return DDQC_E_NULL;
  if (aux.u_ic && gram_cache) return DDQC_E_CONFIG;
  if (!ws || !u_win || !y_win || !y_r || !u_opt || !us_opt || !ys_opt || !status || !iters || !have_prev)
    return DDQC_E_NULL;
  if (batch < 0) return DDQC_E_SHAPE;
  if (batch == 0) return 0;
  const Dims d = make_dims(*cfg);
  const size_t smem = smem_bytes(d);
  cudaError_t e = cudaFuncSetAttribute(ddmpc_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = (int)(batch < sms ? batch : sms);
  if (grid > kMaxSlots) grid = kMaxSlots;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // layout of ws: [work counter, 256 B][slot 0][slot 1]...
  e = cudaMemsetAsync(ws, 0, 256, st);
  if (e != cudaSuccess) return (int)e;
  double* slots = reinterpret_cast<double*>(static_cast<char*>(ws) + 256);
  ddmpc_solve_kernel<<<grid, kThreads, smem, st>>>(*cfg, slots, static_cast<int*>(ws), u_win, y_win, y_r,
                                                   have_prev, lamb, act_state, u_opt, us_opt, ys_opt, status,
                                                   iters, (int)batch, g_phase_clk, static_cast<double*>(gram_cache), aux);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_ddmpc_store(const DdqcMpcConfig* cfg, double* u_win, double* y_win, const double* u_new,
                                const double* y_new, int64_t batch, void* gram_cache, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!u_win || !y_win || !u_new || !y_new) return DDQC_E_NULL;
  if (batch <= 0) return batch == 0 ? 0 : DDQC_E_SHAPE;
  ddmpc_store_kernel<<<(unsigned)batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(*cfg, u_win, y_win, u_new, y_new, (int)batch,
                                                                                    static_cast<double*>(gram_cache));
  return (int)cudaGetLastError();
}

extern "C" int ddqc_ddmpc_action(const DdqcMpcConfig* cfg, const double* u_opt, int32_t n_step, const float* bounds,
                                 float* action, double* u_applied, int64_t batch, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!u_opt || !bounds || !action || !u_applied) return DDQC_E_NULL;
  if (n_step < 0 || n_step > cfg->L || batch < 0) return DDQC_E_SHAPE;
  if (batch == 0) return 0;
  const int tot = (int)batch * cfg->m;
  ddmpc_action_kernel<<<(tot + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      u_opt, cfg->L + 1, cfg->m, n_step, bounds, action, u_applied, (int)batch);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_ddmpc_post_step(const DdqcMpcConfig* cfg, double* u_win, double* y_win, const double* u_applied,
                                    const float* y_meas, int64_t y_rs, int64_t y_cs, const double* y_r,
                                    const double* d0, double max_inc, double* sq_sum, int32_t* n_acc,
                                    int32_t* fail_step, double* y_log, int32_t step, int64_t batch, void* gram_cache,
                                    void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!u_win || !y_win || !u_applied || !y_meas || !y_r || !d0 || !sq_sum || !n_acc || !fail_step) return DDQC_E_NULL;
  if (batch < 0 || step < 0) return DDQC_E_SHAPE;
  if (batch == 0) return 0;
  ddmpc_post_step_kernel<<<(unsigned)batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      *cfg, static_cast<double*>(gram_cache), u_win, y_win, u_applied, y_meas, (long long)y_rs, (long long)y_cs, y_r, d0, max_inc,
      sq_sum, n_acc, fail_step, y_log, step, (int)batch);
  return (int)cudaGetLastError();
}
