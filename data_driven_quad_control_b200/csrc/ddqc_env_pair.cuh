// ddqc_env_pair.cuh -- the env step for TWO envs per thread on packed fp32x2 arithmetic (sm_100a).
//
// The one-thread-per-env kernel is bound by issued instructions, not by HBM: ~1240 instructions per env of which only
// 575 are floating-point arithmetic (no FMA contraction by contract); the rest is 64-bit address arithmetic, constant
// loads, compares / selects and control flow, and ncu shows issue slots at 62-66 % with the DRAM pipe at 60 % (the
// bulk-copy tile kernel, which takes DRAM latency out of the picture altogether, is not faster: DESIGN.md section 3).
// Here a thread owns envs (2j, 2j+1):
//   * every arithmetic operation is ONE packed instruction for both envs - Blackwell's FFMA2 / FADD2
//     (fma.rn.f32x2, add.rn.f32x2, sub.rn.f32x2): IEEE round-to-nearest per lane, so every result is bit-identical to
//     the scalar kernel's and to the oracle's;
//   * every state row is one 64-bit load / store per thread (the two envs are adjacent in the struct-of-arrays rows),
//     i.e. half the memory instructions and half the address arithmetic per env;
//   * the host parameters arrive as packed (c, c) pairs in kernel-parameter space and are used straight from uniform
//     registers (FFMA2 takes a 64-bit uniform operand), loaded once per pair of envs;
//   * compares, selects, square roots and the integer bookkeeping stay per lane.
// A multiplication is issued as fma(a, b, -0.0): ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under
// --fmad=false (checked in SASS), which would break the unfused operation order the reference's torch code has;
// x + (-0) == x for every x, signed zeros included, so the product is exactly round(a * b).
#pragma once
#include <stddef.h>

#include "ddqc_device.cuh"

namespace ddqc {

struct f2 {
  unsigned long long v;  // low half: env 2j, high half: env 2j + 1
};
struct b2 {
  bool a, b;
};

// DdqcEnvParams with every 32-bit word duplicated into a (c, c) pair, plus a few literals
constexpr int kEnvParamWords = (int)(sizeof(DdqcEnvParams) / 4);
enum { L_NEGZERO = 0, L_ZERO, L_ONE, L_TWO, L_HALF, L_ONE_HALF, L_MINUS_HALF, L_MINUS_TWO, L_COUNT };
struct EnvParams2 {
  unsigned long long w[kEnvParamWords];
  unsigned long long lit[L_COUNT];
};
#define K2(field) (f2{PP.w[offsetof(DdqcEnvParams, field) / 4]})
#define K2A(field, j) (f2{PP.w[offsetof(DdqcEnvParams, field) / 4 + (j)]})
#define LIT(name) (f2{PP.lit[name]})

__device__ __forceinline__ f2 pk(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float lo(f2 x) { return __uint_as_float((unsigned int)(x.v & 0xffffffffull)); }
__device__ __forceinline__ float hi(f2 x) { return __uint_as_float((unsigned int)(x.v >> 32)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return d;
}
__device__ __forceinline__ f2 neg2(f2 a) { return f2{a.v ^ 0x8000000080000000ull}; }
template <class FA, class FB>
__device__ __forceinline__ f2 lanes(f2 x, FA fa, FB fb) {
  return pk(fa(lo(x)), fb(hi(x)));
}
__device__ __forceinline__ f2 sel2(b2 m, f2 a, f2 b) { return pk(m.a ? lo(a) : lo(b), m.b ? hi(a) : hi(b)); }
__device__ __forceinline__ f2 clip2(f2 x, float l, float h) { return pk(clipf(lo(x), l, h), clipf(hi(x), l, h)); }
__device__ __forceinline__ f2 sqrt2(f2 x) { return pk(sqrt_nonneg(lo(x)), sqrt_nonneg(hi(x))); }

struct V3p {
  f2 x, y, z;
};
struct Q4p {
  f2 w, x, y, z;
};

// The arithmetic below mirrors ddqc_device.cuh operation for operation (fm -> fma2, * -> mul2, + -> add2, - -> sub2).
struct PairMath {
  const EnvParams2& PP;
  __device__ __forceinline__ f2 mul2(f2 a, f2 b) const { return fma2(a, b, LIT(L_NEGZERO)); }
  // `na` = (-a.x, -a.y, -a.z): the caller holds the conjugate of `a` anyway, so no negation is issued here
  __device__ __forceinline__ Q4p quat_mul(Q4p a, V3p na, Q4p b) const {
    Q4p r;
    r.w = fma2(na.z, b.z, fma2(na.y, b.y, fma2(na.x, b.x, mul2(a.w, b.w))));
    r.x = fma2(na.z, b.y, fma2(a.y, b.z, fma2(a.x, b.w, mul2(a.w, b.x))));
    r.y = fma2(a.z, b.x, fma2(a.y, b.w, fma2(na.x, b.z, mul2(a.w, b.y))));
    r.z = fma2(a.z, b.w, fma2(na.y, b.x, fma2(a.x, b.y, mul2(a.w, b.z))));
    return r;
  }
  // v rotated by q; `nq` = (-q.x, -q.y, -q.z).  -(q.z v.y) of the scalar code is issued as (-q.z) v.y: the same value
  __device__ __forceinline__ V3p rotate_by_quat(V3p v, Q4p q, V3p nq) const {
    f2 ux = fma2(q.y, v.z, mul2(nq.z, v.y));
    f2 uy = fma2(q.z, v.x, mul2(nq.x, v.z));
    f2 uz = fma2(q.x, v.y, mul2(nq.y, v.x));
    f2 sx = fma2(q.y, uz, fma2(nq.z, uy, mul2(q.w, ux)));
    f2 sy = fma2(q.z, ux, fma2(nq.x, uz, mul2(q.w, uy)));
    f2 sz = fma2(q.x, uy, fma2(nq.y, ux, mul2(q.w, uz)));
    return V3p{fma2(LIT(L_TWO), sx, v.x), fma2(LIT(L_TWO), sy, v.y), fma2(LIT(L_TWO), sz, v.z)};
  }
  __device__ __forceinline__ f2 sumsq3(V3p a) const { return add2(add2(mul2(a.x, a.x), mul2(a.y, a.y)), mul2(a.z, a.z)); }
};

struct WrenchP {
  f2 ax, ay, az, cxy, cz;
};

}  // namespace ddqc
