// ddqc_policy.cu -- fused policy-MLP forward (16 -> 128 -> 128 -> A, tanh) on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM) for sm_100a.
//
// Replaces the actor forward of the reference's rollouts: rsl_rl `ActorCritic.act_inference`
// = nn.Sequential(Linear(16,128), Tanh, Linear(128,128), Tanh, Linear(128,A)) with the shape of
// models/demo_ctbr_fixed_yaw/model_1500.pt (learning/config/hover_ppo_config.py:33-38), called once
// per env step on the (N, 16) observation buffer `HoverEnv.step` returns.  With cuBLAS the two
// 128-wide hidden activations make a round trip through HBM (1 KB per env-step, 3x the 357 B of
// the env step itself); here an observation tile goes obs -> smem -> TMEM -> registers -> smem
// -> TMEM -> registers -> action without leaving the SM.
//
// Structure: persistent CTAs (one per SM), two independent 128-thread warpgroups per CTA, each
// walking its own stream of 128-env tiles so that one warpgroup's epilogue (TMEM -> registers,
// bias + tanh, bf16 split, shared-memory stores) overlaps the other's tensor-core work:
//   1. thread r loads observation row r (64 B), splits it into bf16 hi + lo and stores both in the
//      canonical K-major no-swizzle UMMA layout (8 x 16 B core matrices);
//   2. one elected thread issues D1[128x128] = A1 W1^T as three bf16 products
//      (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM) and commits to an mbarrier;
//   3. every thread reads its TMEM lane (tcgen05.ld 32x32b.x32), adds the bias, applies tanh, splits
//      to bf16 hi/lo and writes the layer-2 operand tile;
//   4. D2 = A2 W2^T (8 K-steps x 3 products) issued in two K halves - the first as soon as both thread halves
//      have written their first 32 columns, so it runs under the rest of step 3 - one commit / wait;
//   5. bias + tanh, then the A-wide output layer (A <= 4, 384 FMAs per env) on the CUDA cores
//      straight from the registers; 12-byte action rows are stored coalesced.
// The hi/lo split (two bf16 = 16 mantissa bits per operand, cross term lo*lo dropped, ~1e-5 relative
// per product) keeps the actions within 1e-4 of an fp32 forward (measured 3.2e-5 on the demo
// checkpoint); precision = 1 runs the single hi*hi product (plain bf16 accuracy, one third of the
// tensor-core work).  A three-term split would be fp32-exact but needs 324 KB of operand tiles.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ddqc.h"

namespace {

constexpr int kIn = 16, kHid = 128, kMaxOut = 4, kTileM = 128;
constexpr int kWgThreads = 256, kThreads = 2 * kWgThreads;  // per warpgroup: 2 threads per env row (one per column half)

// shared-memory map (bytes)
constexpr uint32_t kW2Hi = 0, kW2Lo = 32768, kW1Hi = 65536, kW1Lo = 69632;
constexpr uint32_t kWgBase = 73728, kWgStride = 73728;  // per warpgroup: A2 hi | A2 lo | A1 hi | A1 lo
constexpr uint32_t kA2Hi = 0, kA2Lo = 32768, kA1Hi = 65536, kA1Lo = 69632;
constexpr uint32_t kParams = kWgBase + 2 * kWgStride;   // fp32: W3 [4][128] | b1 | b2 | b3 | output partials 2 x [128][4]
constexpr uint32_t kW3 = kParams, kB1 = kW3 + kMaxOut * kHid * 4, kB2 = kB1 + kHid * 4, kB3 = kB2 + kHid * 4;
constexpr uint32_t kPart = kB3 + 64, kPartStride = kTileM * kMaxOut * 4;
constexpr uint32_t kBar = kPart + 2 * kPartStride;      // 2 mbarriers (8 B each) + TMEM base pointer
constexpr uint32_t kSmemBytes = kBar + 32;

// K-major, no swizzle: element (r, k) of an [rows x K] bf16 operand; core matrix = 8 rows x 16 bytes
__device__ __forceinline__ uint32_t core_off(int r, int k, int kchunks) {
  return (uint32_t)((r >> 3) * (kchunks * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start >> 4 in [0,14), leading byte offset
// >> 4 in [16,30) (distance of the two 16-byte K chunks of one MMA), stride byte offset >> 4 in [32,46)
// (distance of consecutive 8-row groups), version 1 in [46,48), layout type 0 = no swizzle in [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D = F32 (1 << 4), A = B = BF16 (1 << 7,
// 1 << 10), both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kHid >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, %1;" ::"r"(wg + 1), "n"(kWgThreads) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive fp32 columns of this thread's TMEM lane, asynchronously: the registers are valid after tmem_wait16(r).
// (tcgen05.wait::ld waits for ALL outstanding loads of the thread, so the pipeline is wait(current) -> issue(next) ->
// arithmetic on current; the "+r" operands of the wait keep the compiler from reading the registers above it.)
#ifndef DDQC_POLICY_TMEM_PIPE
#define DDQC_POLICY_TMEM_PIPE 1
#endif
// DDQC_POLICY_DEFER_OUT = 1: the half-0 thread sums and stores a row's actions after the NEXT warpgroup barrier (under the
// next tile's layer-1 MMAs), one barrier per tile less.  Measured and left off: 0.1749-0.1751 ms either way - that barrier
// is not on the critical path.
// DDQC_POLICY_MMA1_AHEAD = 1: the layer-1 MMAs of tile t+1 are issued right behind the last layer-2 MMAs of tile t (its
// operand tile is stored before the barrier that precedes them), so they run under the layer-2 epilogue of tile t and no
// tile starts by waiting for its own layer-1 product (second mbarrier per warpgroup).
#ifndef DDQC_POLICY_MMA1_AHEAD
#define DDQC_POLICY_MMA1_AHEAD 1
#endif
// DDQC_POLICY_PDL = 1: programmatic dependent launch.  In a rollout the kernel in front of this one is the env step's
// one-warp finalisation and the one behind it the next env step, both PDL-aware (ddqc_env.cu): launched with the
// programmatic-serialisation attribute, this kernel converts its weights / allocates TMEM while the finalisation still
// runs, waits (griddepcontrol.wait) before it touches the observations, and only then releases its own dependents - the
// next env step becomes resident on the SMs whose tile streams have ended and issues its state loads under the tail.
#ifndef DDQC_POLICY_PDL
#define DDQC_POLICY_PDL 1
#endif
#ifndef DDQC_POLICY_DEFER_OUT
#define DDQC_POLICY_DEFER_OUT 0
#endif
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// x = hi + lo with hi, lo bf16 (16 mantissa bits together); two values per conversion instruction
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);  // .x = a in the low half
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 8 consecutive K elements -> one 16-byte chunk of the hi operand and one of the lo operand
__device__ __forceinline__ void store_chunk(uint8_t* smem, uint32_t off_hi, uint32_t off_lo, const float (&x)[8]) {
  uint4 h, l;
  split_pair(x[0], x[1], h.x, l.x);
  split_pair(x[2], x[3], h.y, l.y);
  split_pair(x[4], x[5], h.z, l.z);
  split_pair(x[6], x[7], h.w, l.w);
  *reinterpret_cast<uint4*>(smem + off_hi) = h;
  *reinterpret_cast<uint4*>(smem + off_lo) = l;
}

// tanh(x) = 1 - 2 / (1 + exp(2x)): ex2.approx + rcp.approx, absolute error ~1e-7; saturates cleanly (exp -> inf gives 1,
// exp -> 0 gives -1), so no clamp is needed.  Callers pass xs = 2 log2(e) x: the scale is folded into the bias add (the
// biases are stored pre-scaled), one FFMA.  Measured and dropped: the exponential of 2-4 of every 4 elements on the FMA
// pipe (round-to-nearest split + degree-4 minimax 2^f + exponent insert) to unload the MUFU pipe (ncu: XU 53 % busy,
// the busiest unit) - 0.225 / 0.239 / 0.250 ms against 0.217 ms: the epilogue is bound by issued instructions per
// tile, not by the MUFU rate, and each polynomial exponential costs ~9 more of them.
constexpr float kTanhScale = 2.885390081777927f;
[[maybe_unused]] __device__ __forceinline__ float tanh_acc(float xs) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xs));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(-2.0f, r, 1.0f);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2: two fp32 results per issued instruction) ----------------
// The epilogue is bound by issued instructions (4 warps per scheduler), so everything in it that is not a MUFU or a
// conversion runs two elements per instruction: the tanh argument, 1 + e, 1 - 2r, the bf16 split residual and the
// output-layer accumulation.
#ifndef DDQC_POLICY_F32X2
#define DDQC_POLICY_F32X2 1
#endif
#ifndef DDQC_POLICY_KSPLIT
#define DDQC_POLICY_KSPLIT 1  // layer-2 MMAs issued in two K halves, the first under the layer-1 epilogue (0.2002 vs 0.2035 ms)
#endif
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// tanh of two pre-scaled arguments
__device__ __forceinline__ f2_t tanh2(f2_t xs) {
  float x0, x1, e0, e1, d0, d1, r0, r1;
  upk2(xs, x0, x1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
  upk2(add2(pk2(e0, e1), pk2(1.0f, 1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return fma2(pk2(-2.0f, -2.0f), pk2(r0, r1), pk2(1.0f, 1.0f));
}
// tanh of FOUR pre-scaled arguments with ONE reciprocal: with d_i = 1 + 2^x_i, A = (d0, d1), B = (d2, d3), p = A * B,
// r = 1 / (p.lo p.hi) and s = r (p.hi, p.lo) = (1 / (d0 d2), 1 / (d1 d3)), the four reciprocals are s * B and s * A
// (packed), so four tanh cost 5 MUFU (4 ex2 + 1 rcp) instead of 8 - the XU pipe is the busiest unit of this kernel
// (ncu: 41 % of the stall samples sit on MUFU instructions).  The arguments are clamped to 30 first: 1 - 2 / (1 + 2^30)
// already rounds to 1.0f, and the product of four denominators stays below 2^121.  Measured: 0.182 ms per 2^20 envs
// against 0.199 ms with two MUFU per tanh, same accuracy (3.8e-5 / 3.6e-5 on the reference checkpoint).
#ifndef DDQC_POLICY_TANH4
#define DDQC_POLICY_TANH4 1
#endif
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void tanh4(f2_t xa, f2_t xb, f2_t& ta, f2_t& tb) {
#if DDQC_POLICY_TANH4
  float x0, x1, x2, x3, e0, e1, e2, e3, p0, p1, r;
  upk2(xa, x0, x1);
  upk2(xb, x2, x3);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(x0, 30.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(x1, 30.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fminf(x2, 30.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fminf(x3, 30.0f)));
  const f2_t one = pk2(1.0f, 1.0f);
  const f2_t A = add2(pk2(e0, e1), one), B = add2(pk2(e2, e3), one);
  upk2(mul2(A, B), p0, p1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p0 * p1));
  const f2_t sw = mul2(pk2(r, r), pk2(p1, p0));
  const f2_t m2 = pk2(-2.0f, -2.0f);
  ta = fma2(m2, mul2(sw, B), one);
  tb = fma2(m2, mul2(sw, A), one);
#else
  ta = tanh2(xa);
  tb = tanh2(xb);
#endif
}
// bf16 hi/lo split of a packed pair; the residual y - hi is one FFMA2
__device__ __forceinline__ void split_pair2(f2_t y, uint32_t& hi, uint32_t& lo) {
  float a, b, ra, rb;
  upk2(y, a, b);
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  upk2(fma2(pk2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u)), pk2(-1.0f, -1.0f), y), ra, rb);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
policy_mlp_kernel(const float* __restrict__ obs, const long long n, const float* __restrict__ W1,
                  const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                  const float* __restrict__ W3, const float* __restrict__ b3, const int n_out,
                  float* __restrict__ actions) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, wg = tid >> 8, wtid = tid & 255, warp = tid >> 5;
  const int row = wtid & 127, half = wtid >> 7;  // env row of the tile / column half this thread owns
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kBar + 16);

  // ---- one-time setup: weights -> bf16 hi/lo operand tiles, fp32 parameters, barriers, TMEM ----
  for (int e = tid; e < kHid * kHid / 8; e += kThreads) {  // W2[n][k]: 8 consecutive k per thread
    const int nn = e / (kHid / 8), k0 = (e % (kHid / 8)) * 8;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = W2[nn * kHid + k0 + i];
    const uint32_t off = core_off(nn, k0, kHid / 8);
    store_chunk(smem, kW2Hi + off, kW2Lo + off, x);
  }
  for (int e = tid; e < kHid * kIn / 8; e += kThreads) {
    const int nn = e / (kIn / 8), k0 = (e % (kIn / 8)) * 8;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = W1[nn * kIn + k0 + i];
    const uint32_t off = core_off(nn, k0, kIn / 8);
    store_chunk(smem, kW1Hi + off, kW1Lo + off, x);
  }
  float* sW3 = reinterpret_cast<float*>(smem + kW3);
  float* sB1 = reinterpret_cast<float*>(smem + kB1);
  float* sB2 = reinterpret_cast<float*>(smem + kB2);
  float* sB3 = reinterpret_cast<float*>(smem + kB3);
  for (int e = tid; e < kMaxOut * kHid; e += kThreads) sW3[e] = e < n_out * kHid ? W3[e] : 0.0f;
  for (int e = tid; e < kHid; e += kThreads) {
    sB1[e] = b1[e] * kTanhScale;  // pre-scaled: tanh argument = fma(acc, kTanhScale, bias)
    sB2[e] = b2[e] * kTanhScale;
  }
  if (tid < kMaxOut) sB3[tid] = tid < n_out ? b3[tid] : 0.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbase + kBar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbase + kBar + 8));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbase + kB3 + 16));  // layer-1 barriers (MMA1_AHEAD): the
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbase + kB3 + 24));  // unused tail of the b3 slot
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kBar + 16), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t d1 = tmem_base + wg * 256, d2 = d1 + 128;          // accumulator columns of this warpgroup
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;     // this warp's 32-lane slice of TMEM (= row / 32)
  const uint32_t mbar = sbase + kBar + 8 * wg;
  uint8_t* wgs = smem + kWgBase + wg * kWgStride;
  const uint32_t wga = sbase + kWgBase + wg * kWgStride;
  float* part = reinterpret_cast<float*>(smem + kPart + wg * kPartStride);
  uint32_t phase = 0;
  constexpr bool kAhead = DDQC_POLICY_MMA1_AHEAD && DDQC_POLICY_TMEM_PIPE && DDQC_POLICY_F32X2 && DDQC_POLICY_KSPLIT;
  const uint32_t mbar1 = kAhead ? sbase + kB3 + 16 + 8 * wg : mbar;  // layer-1 MMAs complete
  uint32_t phase1 = 0;

  const long long n_tiles = (n + kTileM - 1) / kTileM;
  const long long tile_step = (long long)gridDim.x * 2;
#if DDQC_POLICY_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");               // the producer of `obs` has completed
#if DDQC_POLICY_PDL == 2
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the consumer of `actions` may be scheduled
#endif
#endif
  // observation chunk of this thread (row, 8 of the 16 inputs), software-pipelined: the loads of the NEXT tile are
  // issued while the layer-2 MMAs of the current one run, so no tile starts with an exposed global-memory round trip
  float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
  {
    const long long t0 = (long long)blockIdx.x * 2 + wg;
    if (t0 < n_tiles && t0 * kTileM + row < n) {
      const float4* src = reinterpret_cast<const float4*>(obs + (t0 * kTileM + row) * kIn + 8 * half);
      o0 = __ldg(src);
      o1 = __ldg(src + 1);
    }
  }
#if DDQC_POLICY_DEFER_OUT
  float pend_out[kMaxOut] = {0.0f, 0.0f, 0.0f, 0.0f};
  long long pend_row = -1;
  auto flush_pending = [&]() {
    if (half == 0 && pend_row >= 0) {
      const float4 o = *reinterpret_cast<const float4*>(part + row * kMaxOut);
      const float po[kMaxOut] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int j = 0; j < kMaxOut; ++j)
        if (j < n_out) actions[pend_row * n_out + j] = (pend_out[j] + po[j]) + sB3[j];
      pend_row = -1;
    }
  };
#endif
  // ---- 1. observation row (registers o0, o1) -> layer-1 operand tile (each thread one 8-element K chunk) ----
  auto store_a1 = [&]() {
    float y[8];
    y[0] = o0.x; y[1] = o0.y; y[2] = o0.z; y[3] = o0.w; y[4] = o1.x; y[5] = o1.y; y[6] = o1.z; y[7] = o1.w;
    const uint32_t off = core_off(row, 8 * half, kIn / 8);
    store_chunk(wgs, kA1Hi + off, kA1Lo + off, y);
  };
  // ---- 2. D1 = A1 W1^T (one elected thread; behind a warpgroup barrier that follows the operand stores) ----
  auto issue_mma1 = [&]() {
    const uint64_t ah = make_desc(wga + kA1Hi, 128, (kIn / 8) * 128), al = make_desc(wga + kA1Lo, 128, (kIn / 8) * 128);
    const uint64_t bh = make_desc(sbase + kW1Hi, 128, (kIn / 8) * 128), bl = make_desc(sbase + kW1Lo, 128, (kIn / 8) * 128);
    umma_bf16(d1, ah, bh, 0);
    if (SPLIT) {
      umma_bf16(d1, al, bh, 1);
      umma_bf16(d1, ah, bl, 1);
    }
    umma_commit(mbar1);
  };
  auto load_obs = [&](long long t) {  // observation chunk of tile t -> o0, o1 (zeros past the end)
    const long long r = t * kTileM + row;
    o0 = make_float4(0.f, 0.f, 0.f, 0.f);
    o1 = o0;
    if (t < n_tiles && r < n) {
      const float4* src = reinterpret_cast<const float4*>(obs + r * kIn + 8 * half);
      o0 = __ldg(src);
      o1 = __ldg(src + 1);
    }
  };
  if (kAhead && (long long)blockIdx.x * 2 + wg < n_tiles) {  // prologue: the first tile's layer-1 product
    store_a1();
    fence_async_smem();
    tc_fence_before();
    wg_sync(wg);
    if (wtid == 0) {
      tc_fence_after();
      issue_mma1();
    }
    load_obs((long long)blockIdx.x * 2 + wg + tile_step);
  }
  for (long long tile = (long long)blockIdx.x * 2 + wg; tile < n_tiles; tile += tile_step) {
    const long long grow = tile * kTileM + row;
    const bool valid = grow < n;
    const bool has_next = tile + tile_step < n_tiles;
    if (!kAhead) {
      store_a1();
      fence_async_smem();
      tc_fence_before();
      wg_sync(wg);
      if (wtid == 0) {
        tc_fence_after();
        issue_mma1();
      }
    }
#if DDQC_POLICY_DEFER_OUT
    flush_pending();  // the previous tile's actions, under the layer-1 MMAs
#endif
    mbar_wait(mbar1, kAhead ? phase1 : phase);
    if (kAhead) phase1 ^= 1; else phase ^= 1;
    tc_fence_after();
#if DDQC_POLICY_TMEM_PIPE && DDQC_POLICY_F32X2 && DDQC_POLICY_KSPLIT
    // ---- 3 + 4, TMEM reads software-pipelined: 16 columns at a time, the load of the next 16 in flight under the
    // arithmetic of the current ones (ncu: 20 % of the stall cycles of the one-shot x32 version were spent behind
    // tcgen05.wait::ld); the layer-2 MMAs of a 32-column pair of K steps are issued as soon as it is written, as before
    {
      uint32_t ra[16], rb[16];
      tmem_ld16_issue(d1 + lane_addr + half * 64, ra);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t(&cur)[16] = (q & 1) ? rb : ra;
        uint32_t(&nxt)[16] = (q & 1) ? ra : rb;
        const int c0 = half * 64 + q * 16;
        tmem_wait16(cur);
        if (q < 3) tmem_ld16_issue(d1 + lane_addr + c0 + 16, nxt);
        const f2_t sc = pk2(kTanhScale, kTanhScale);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 ba = *reinterpret_cast<const float4*>(sB1 + c0 + 8 * c), bb = *reinterpret_cast<const float4*>(sB1 + c0 + 8 * c + 4);
          uint4 h, l;
          f2_t t0, t1, t2, t3;
          tanh4(fma2(pk2(__uint_as_float(cur[8 * c]), __uint_as_float(cur[8 * c + 1])), sc, pk2(ba.x, ba.y)),
                fma2(pk2(__uint_as_float(cur[8 * c + 2]), __uint_as_float(cur[8 * c + 3])), sc, pk2(ba.z, ba.w)), t0, t1);
          tanh4(fma2(pk2(__uint_as_float(cur[8 * c + 4]), __uint_as_float(cur[8 * c + 5])), sc, pk2(bb.x, bb.y)),
                fma2(pk2(__uint_as_float(cur[8 * c + 6]), __uint_as_float(cur[8 * c + 7])), sc, pk2(bb.z, bb.w)), t2, t3);
          split_pair2(t0, h.x, l.x);
          split_pair2(t1, h.y, l.y);
          split_pair2(t2, h.z, l.z);
          split_pair2(t3, h.w, l.w);
          const uint32_t off = core_off(row, c0 + 8 * c, kHid / 8);
          *reinterpret_cast<uint4*>(wgs + kA2Hi + off) = h;
          *reinterpret_cast<uint4*>(wgs + kA2Lo + off) = l;
        }
        if (q & 1) {  // the 32-column pair cb = q / 2 of both thread halves is complete: contract it
          const int cb = q >> 1;
          // MMA1_AHEAD: D1 has been read completely and the layer-1 MMAs of this tile are done (waited for above), so
          // the NEXT tile's layer-1 operand tile can be stored now and contracted right behind this tile's layer 2
          if (kAhead && cb == 1 && has_next) store_a1();
          fence_async_smem();
          tc_fence_before();
          wg_sync(wg);
          if (wtid == 0) {
            tc_fence_after();
#pragma unroll 1
            for (int kq = 0; kq < 4; ++kq) {
              const int ks = (kq >> 1) * 4 + cb * 2 + (kq & 1);
              const uint32_t ko = ks * 256;  // two 16-byte K chunks per MMA
              const uint64_t ah = make_desc(wga + kA2Hi + ko, 128, (kHid / 8) * 128), al = make_desc(wga + kA2Lo + ko, 128, (kHid / 8) * 128);
              const uint64_t bh = make_desc(sbase + kW2Hi + ko, 128, (kHid / 8) * 128), bl = make_desc(sbase + kW2Lo + ko, 128, (kHid / 8) * 128);
              umma_bf16(d2, ah, bh, (cb | kq) != 0);
              if (SPLIT) {
                umma_bf16(d2, al, bh, 1);
                umma_bf16(d2, ah, bl, 1);
              }
            }
            if (cb == 1) umma_commit(mbar);
            if (kAhead && cb == 1 && has_next) issue_mma1();
          }
        }
      }
    }
#else
    // ---- 3. h1 = tanh(D1 + b1) -> layer-2 operand tile (this thread: 64 of the row's 128 columns) ----
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int c0 = half * 64 + cb * 32;
      float v[32];
      tmem_ld32(d1 + lane_addr + c0, v);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 ba = *reinterpret_cast<const float4*>(sB1 + c0 + 8 * c), bb = *reinterpret_cast<const float4*>(sB1 + c0 + 8 * c + 4);
#if DDQC_POLICY_F32X2
        const f2_t sc = pk2(kTanhScale, kTanhScale);
        uint4 h, l;
        f2_t t0, t1, t2, t3;
        tanh4(fma2(pk2(v[8 * c], v[8 * c + 1]), sc, pk2(ba.x, ba.y)), fma2(pk2(v[8 * c + 2], v[8 * c + 3]), sc, pk2(ba.z, ba.w)), t0, t1);
        tanh4(fma2(pk2(v[8 * c + 4], v[8 * c + 5]), sc, pk2(bb.x, bb.y)), fma2(pk2(v[8 * c + 6], v[8 * c + 7]), sc, pk2(bb.z, bb.w)), t2, t3);
        split_pair2(t0, h.x, l.x);
        split_pair2(t1, h.y, l.y);
        split_pair2(t2, h.z, l.z);
        split_pair2(t3, h.w, l.w);
        const uint32_t off = core_off(row, c0 + 8 * c, kHid / 8);
        *reinterpret_cast<uint4*>(wgs + kA2Hi + off) = h;
        *reinterpret_cast<uint4*>(wgs + kA2Lo + off) = l;
#else
        float y[8];
        y[0] = tanh_acc(fmaf(v[8 * c], kTanhScale, ba.x)); y[1] = tanh_acc(fmaf(v[8 * c + 1], kTanhScale, ba.y));
        y[2] = tanh_acc(fmaf(v[8 * c + 2], kTanhScale, ba.z)); y[3] = tanh_acc(fmaf(v[8 * c + 3], kTanhScale, ba.w));
        y[4] = tanh_acc(fmaf(v[8 * c + 4], kTanhScale, bb.x)); y[5] = tanh_acc(fmaf(v[8 * c + 5], kTanhScale, bb.y));
        y[6] = tanh_acc(fmaf(v[8 * c + 6], kTanhScale, bb.z)); y[7] = tanh_acc(fmaf(v[8 * c + 7], kTanhScale, bb.w));
        const uint32_t off = core_off(row, c0 + 8 * c, kHid / 8);
        store_chunk(wgs, kA2Hi + off, kA2Lo + off, y);
#endif
      }
      // ---- 4. D2 = A2 W2^T, K split in two: the 64 columns both thread halves have just written (cb*32 .. +32 and
      // 64 + cb*32 .. +32) are contracted while the warpgroup computes its other 64 columns, so half of the layer-2
      // tensor work is hidden under the layer-1 epilogue instead of being waited for
      if (DDQC_POLICY_KSPLIT || cb == 1) {
        fence_async_smem();
        tc_fence_before();
        wg_sync(wg);
        if (wtid == 0) {
          tc_fence_after();
#pragma unroll 1
          for (int kq = 0; kq < (DDQC_POLICY_KSPLIT ? 4 : 8); ++kq) {
            const int ks = DDQC_POLICY_KSPLIT ? (kq >> 1) * 4 + cb * 2 + (kq & 1) : kq;
            const uint32_t ko = ks * 256;  // two 16-byte K chunks per MMA
            const uint64_t ah = make_desc(wga + kA2Hi + ko, 128, (kHid / 8) * 128), al = make_desc(wga + kA2Lo + ko, 128, (kHid / 8) * 128);
            const uint64_t bh = make_desc(sbase + kW2Hi + ko, 128, (kHid / 8) * 128), bl = make_desc(sbase + kW2Lo + ko, 128, (kHid / 8) * 128);
            umma_bf16(d2, ah, bh, DDQC_POLICY_KSPLIT ? (cb | kq) != 0 : kq > 0);
            if (SPLIT) {
              umma_bf16(d2, al, bh, 1);
              umma_bf16(d2, ah, bl, 1);
            }
          }
          if (cb == 1) umma_commit(mbar);
        }
      }
    }
#endif
    load_obs(tile + (kAhead ? 2 : 1) * tile_step);  // prefetch: the next tile whose operand tile is not stored yet
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- 5. h2 = tanh(D2 + b2); action = h2 W3^T + b3 on the CUDA cores (partial over this column half) ----
    float out[kMaxOut];
#if DDQC_POLICY_F32X2
    f2_t out2[kMaxOut];  // (even, odd) column partial sums
#pragma unroll
    for (int j = 0; j < kMaxOut; ++j) out2[j] = pk2(0.0f, 0.0f);
#if DDQC_POLICY_TMEM_PIPE
    {
      uint32_t ra[16], rb[16];
      tmem_ld16_issue(d2 + lane_addr + half * 64, ra);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t(&cur)[16] = (q & 1) ? rb : ra;
        uint32_t(&nxt)[16] = (q & 1) ? ra : rb;
        const int c0 = half * 64 + q * 16;
        tmem_wait16(cur);
        if (q < 3) tmem_ld16_issue(d2 + lane_addr + c0 + 16, nxt);
        const f2_t sc = pk2(kTanhScale, kTanhScale);
        f2_t h[8];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float4 bq = *reinterpret_cast<const float4*>(sB2 + c0 + 4 * qq);
          tanh4(fma2(pk2(__uint_as_float(cur[4 * qq]), __uint_as_float(cur[4 * qq + 1])), sc, pk2(bq.x, bq.y)),
                fma2(pk2(__uint_as_float(cur[4 * qq + 2]), __uint_as_float(cur[4 * qq + 3])), sc, pk2(bq.z, bq.w)), h[2 * qq], h[2 * qq + 1]);
        }
#pragma unroll
        for (int j = 0; j < kMaxOut; ++j) {
          if (j < n_out) {  // uniform
            const float4* w = reinterpret_cast<const float4*>(sW3 + j * kHid + c0);
            f2_t acc = out2[j];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const float4 ww = w[qq];
              acc = fma2(h[2 * qq], pk2(ww.x, ww.y), acc);
              acc = fma2(h[2 * qq + 1], pk2(ww.z, ww.w), acc);
            }
            out2[j] = acc;
          }
        }
      }
    }
#else
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int c0 = half * 64 + cb * 32;
      float v[32];
      tmem_ld32(d2 + lane_addr + c0, v);
      const f2_t sc = pk2(kTanhScale, kTanhScale);
      f2_t h[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bq = *reinterpret_cast<const float4*>(sB2 + c0 + 4 * q);
        tanh4(fma2(pk2(v[4 * q], v[4 * q + 1]), sc, pk2(bq.x, bq.y)), fma2(pk2(v[4 * q + 2], v[4 * q + 3]), sc, pk2(bq.z, bq.w)), h[2 * q],
              h[2 * q + 1]);
      }
#pragma unroll
      for (int j = 0; j < kMaxOut; ++j) {
        if (j < n_out) {  // uniform
          const float4* w = reinterpret_cast<const float4*>(sW3 + j * kHid + c0);
          f2_t acc = out2[j];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 ww = w[q];
            acc = fma2(h[2 * q], pk2(ww.x, ww.y), acc);
            acc = fma2(h[2 * q + 1], pk2(ww.z, ww.w), acc);
          }
          out2[j] = acc;
        }
      }
    }
#endif
#pragma unroll
    for (int j = 0; j < kMaxOut; ++j) {
      float e, o;
      upk2(out2[j], e, o);
      out[j] = e + o;
    }
#else
#pragma unroll
    for (int j = 0; j < kMaxOut; ++j) out[j] = 0.0f;
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      const int c0 = half * 64 + cb * 32;
      float v[32];
      tmem_ld32(d2 + lane_addr + c0, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 bq = *reinterpret_cast<const float4*>(sB2 + c0 + 4 * q);
        v[4 * q] = tanh_acc(fmaf(v[4 * q], kTanhScale, bq.x));
        v[4 * q + 1] = tanh_acc(fmaf(v[4 * q + 1], kTanhScale, bq.y));
        v[4 * q + 2] = tanh_acc(fmaf(v[4 * q + 2], kTanhScale, bq.z));
        v[4 * q + 3] = tanh_acc(fmaf(v[4 * q + 3], kTanhScale, bq.w));
      }
#pragma unroll
      for (int j = 0; j < kMaxOut; ++j) {
        if (j < n_out) {  // uniform
          const float4* w = reinterpret_cast<const float4*>(sW3 + j * kHid + c0);
          float acc = out[j];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 ww = w[q];
            acc = fmaf(v[4 * q], ww.x, acc);
            acc = fmaf(v[4 * q + 1], ww.y, acc);
            acc = fmaf(v[4 * q + 2], ww.z, acc);
            acc = fmaf(v[4 * q + 3], ww.w, acc);
          }
          out[j] = acc;
        }
      }
    }
#endif
    tc_fence_before();  // the TMEM reads above are ordered before the next tile's MMAs (via the next wg_sync)
    if (half == 1) *reinterpret_cast<float4*>(part + row * kMaxOut) = make_float4(out[0], out[1], out[2], out[3]);
#if DDQC_POLICY_DEFER_OUT
    // the two column halves of a row are summed and stored by the half-0 thread AFTER the next warpgroup barrier (the one
    // in front of the next tile's layer-1 MMAs, or the one behind the loop): one barrier per tile less, and the store
    // happens while that MMA runs
    if (half == 0) {
#pragma unroll
      for (int j = 0; j < kMaxOut; ++j) pend_out[j] = out[j];
      pend_row = valid ? grow : -1;
    }
  }
  wg_sync(wg);
  flush_pending();
#else
    wg_sync(wg);
    if (half == 0 && valid) {
      const float4 o = *reinterpret_cast<const float4*>(part + row * kMaxOut);
      const float po[kMaxOut] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int j = 0; j < kMaxOut; ++j)
        if (j < n_out) actions[grow * n_out + j] = (out[j] + po[j]) + sB3[j];
    }
  }
#endif

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

}  // namespace

extern "C" int ddqc_policy_mlp_forward(const float* obs, int64_t n, int32_t n_in, int32_t n_hidden, int32_t n_out,
                                       const float* W1, const float* b1, const float* W2, const float* b2,
                                       const float* W3, const float* b3, int32_t precision, float* actions,
                                       void* stream) {
  if (!obs || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !actions) return DDQC_E_NULL;
  if (n_in != kIn || n_hidden != kHid || n_out < 1 || n_out > kMaxOut || n < 0) return DDQC_E_SHAPE;
  if (precision != 0 && precision != 1) return DDQC_E_CONFIG;
  if ((reinterpret_cast<uintptr_t>(obs) & 15) != 0) return DDQC_E_ALIGN;
  if (n == 0) return 0;
  auto kern = precision == 0 ? policy_mlp_kernel<true> : policy_mlp_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (n + kTileM - 1) / kTileM;
  const int grid = (int)((n_tiles + 1) / 2 < sms ? (n_tiles + 1) / 2 : sms);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cfg.attrs = attr;
  cfg.numAttrs = DDQC_POLICY_PDL ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, obs, (long long)n, W1, b1, W2, b2, W3, b3, (int)n_out, actions);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}
