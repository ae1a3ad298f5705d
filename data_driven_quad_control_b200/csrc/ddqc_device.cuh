// ddqc_device.cuh -- device helpers shared by the env / controller kernels.
// Compiled with -fmad=false: every expression is evaluated exactly as written (see ddqc_env.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ddqc.h"

namespace ddqc {

struct Q4 {
  float w, x, y, z;
};
struct V3 {
  float x, y, z;
};

// fm(a,b,c) = round(a*b + c): the only fused operation in these translation units
// (-fmad=false keeps every other * and + separately rounded, as written).
__device__ __forceinline__ float fm(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Hamilton product a (x) b -- oracle/drone_physics.py: quat_mul
__device__ __forceinline__ Q4 quat_mul(Q4 a, Q4 b) {
  Q4 r;
  r.w = fm(-a.z, b.z, fm(-a.y, b.y, fm(-a.x, b.x, a.w * b.w)));
  r.x = fm(-a.z, b.y, fm(a.y, b.z, fm(a.x, b.w, a.w * b.x)));
  r.y = fm(a.z, b.x, fm(a.y, b.w, fm(-a.x, b.z, a.w * b.y)));
  r.z = fm(a.z, b.w, fm(-a.y, b.x, fm(a.x, b.y, a.w * b.z)));
  return r;
}

// genesis.utils.geom.transform_by_quat: v + 2 (w u + qv x u), u = qv x v -- oracle: rotate_by_quat
__device__ __forceinline__ V3 rotate_by_quat(V3 v, Q4 q) {
  float ux = fm(q.y, v.z, -(q.z * v.y));
  float uy = fm(q.z, v.x, -(q.x * v.z));
  float uz = fm(q.x, v.y, -(q.y * v.x));
  float sx = fm(q.y, uz, fm(-q.z, uy, q.w * ux));
  float sy = fm(q.z, ux, fm(-q.x, uz, q.w * uy));
  float sz = fm(q.x, uy, fm(-q.y, ux, q.w * uz));
  return V3{fm(2.0f, sx, v.x), fm(2.0f, sy, v.y), fm(2.0f, sz, v.z)};
}

__device__ __forceinline__ float horner5(const float* c, float s) {
  float acc = fm(s, c[4], c[3]);
  acc = fm(acc, s, c[2]);
  acc = fm(acc, s, c[1]);
  return fm(acc, s, c[0]);
}

// sqrtf(x) for x >= +0, correctly rounded like sqrtf, but arranged so that the common inputs
// never leave the inline MUFU.RSQ fast path: exact zeros (clamped rotor thrusts, reset envs) and
// the practically unreachable range (0, 1e-30) would otherwise drag the whole warp through the
// out-of-line special-case routine of sqrt.rn.f32.
// Measured and dropped (round 2): the fast path written out as 8 straight-line instructions (max.NaN, MUFU.RSQ, two
// FMUL, two FFMA, select) - ~8 instructions fewer per call, but the compiler then overlaps the four rotor square roots,
// the 64-register step kernel spills 32 bytes per thread and a 2^20-env step takes 67.8 us instead of 63.6 us.
__device__ __forceinline__ float sqrt_nonneg(float x) {
  float r = sqrtf(fmaxf(x, 1.0e-30f));
  if (x < 1.0e-30f) r = (x == 0.0f) ? 0.0f : sqrtf(x);
  return r;
}

// torch.clip(x, lo, hi): NaN passes through (both compares are false).  Measured and dropped (round 2): the
// two-instruction form max.NaN / min.NaN (FMNMX.NAN: same values, NaN included) - 26 instructions fewer per env-step, but
// the 64-register step kernel then spills 4 bytes per thread and a 2^20-env step takes 64.1 us instead of 63.5 us.
__device__ __forceinline__ float clipf(float x, float lo, float hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}

// ---- Philox-4x32-10 -----------------------------------------------------------------
struct U4 {
  uint32_t a, b, c, d;
};
__device__ __forceinline__ U4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                     uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}
__device__ __forceinline__ float u2f(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// Box-Muller pair from two words: (0,1] x [0,1)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  float u2 = u2f(b);
  float rad = sqrtf(-2.0f * logf(u1));
  float ang = 6.283185307179586f * u2;
  z0 = rad * cosf(ang);
  z1 = rad * sinf(ang);
}

// `count` <= 16 standard normals for env `i` from stream `stream`
template <int COUNT>
__device__ __forceinline__ void philox_normals(uint32_t i, uint64_t ev, uint32_t stream, uint64_t seed,
                                               float* z) {
#pragma unroll
  for (int sub = 0; sub * 4 < COUNT; ++sub) {
    U4 r = philox(i, (uint32_t)ev, (uint32_t)(ev >> 32), stream | ((uint32_t)sub << 8), (uint32_t)seed,
                  (uint32_t)(seed >> 32));
    float a, b, c, d;
    box_muller(r.a, r.b, a, b);
    box_muller(r.c, r.d, c, d);
    if (sub * 4 + 0 < COUNT) z[sub * 4 + 0] = a;
    if (sub * 4 + 1 < COUNT) z[sub * 4 + 1] = b;
    if (sub * 4 + 2 < COUNT) z[sub * 4 + 2] = c;
    if (sub * 4 + 3 < COUNT) z[sub * 4 + 3] = d;
  }
}

__device__ __forceinline__ V3 resample_command(const DdqcEnvParams& P, uint32_t i, uint64_t ev, uint32_t stream) {
  U4 r = philox(i, (uint32_t)ev, (uint32_t)(ev >> 32), stream, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
  V3 c;
  c.x = P.cmd_span[0] * u2f(r.a) + P.cmd_lo[0];  // gs_rand_float: (upper-lower)*rand + lower
  c.y = P.cmd_span[1] * u2f(r.b) + P.cmd_lo[1];
  c.z = P.cmd_span[2] * u2f(r.c) + P.cmd_lo[2];
  return c;
}

// ---- rigid body (oracle/drone_physics.py: rotor_wrench, substep) ---------------------------
struct Wrench {
  float ax, ay, az;  // per-substep body-rate increments from the rotor torques
  float cxy, cz;     // 2 dt F/m and dt F/m
};

__device__ __forceinline__ Wrench rotor_wrench(const DdqcEnvParams& P, const float rpm[4]) {
  float s0 = rpm[0] * rpm[0], s1 = rpm[1] * rpm[1], s2 = rpm[2] * rpm[2], s3 = rpm[3] * rpm[3];
  float t0 = s0 * P.kf, t1 = s1 * P.kf, t2 = s2 * P.kf, t3 = s3 * P.kf;
  float fz = ((t0 + t1) + t2) + t3;
  float tau_x = fm(t3, P.arm_x[3], fm(t2, P.arm_x[2], fm(t1, P.arm_x[1], t0 * P.arm_x[0])));
  float tau_y = fm(t3, P.arm_y[3], fm(t2, P.arm_y[2], fm(t1, P.arm_y[1], t0 * P.arm_y[0])));
  float tau_z = fm(s3, P.km_spin[3], fm(s2, P.km_spin[2], fm(s1, P.km_spin[1], s0 * P.km_spin[0])));
  return Wrench{P.k_w[0] * tau_x, P.k_w[1] * tau_y, P.k_w[2] * tau_z, P.two_dt_over_m * fz, P.dt_over_m * fz};
}

// One substep is: pose with the OLD velocities, then the velocities with the forces at the NEW pose (order pinned by
// the reference's published Genesis run: oracle/drone_physics.py).  It is evaluated in two halves so
// that the decimation loop can defer the linear-velocity half of substep k to the top of substep k + 1, where it runs
// beside the (long, dependent) quaternion chain instead of behind it - the same operations on the same operands, only
// their place in the instruction stream differs (bit-identical to substep()).
__device__ __forceinline__ void substep_velocity(const DdqcEnvParams& P, const Wrench& W, const Q4& q, V3& v) {
  // world-frame linear velocity: third column of R(q) at the new pose * thrust / m + gravity
  v.x = fm(W.cxy, fm(q.x, q.z, q.w * q.y), v.x);
  v.y = fm(W.cxy, fm(q.y, q.z, -(q.w * q.x)), v.y);
  float r33 = fm(-2.0f, fm(q.x, q.x, q.y * q.y), 1.0f);
  v.z = fm(W.cz, r33, v.z - P.dt_g);
}

__device__ __forceinline__ void substep_pose_rates(const DdqcEnvParams& P, const Wrench& W, V3& p, Q4& q, const V3& v, V3& w) {
  p.x = fm(v.x, P.dt, p.x);
  p.y = fm(v.y, P.dt, p.y);
  p.z = fm(v.z, P.dt, p.z);
  // q <- renormalise(q (x) [cos|h|, sinc|h| h]), h = w dt/2
  float hx = w.x * P.half_dt, hy = w.y * P.half_dt, hz = w.z * P.half_dt;
  float s = fm(hz, hz, fm(hy, hy, hx * hx));
  float c = horner5(P.cos_coef, s);
  float k = horner5(P.sinc_coef, s);
  Q4 n = quat_mul(q, Q4{c, k * hx, k * hy, k * hz});
  float n2 = fm(n.z, n.z, fm(n.y, n.y, fm(n.x, n.x, n.w * n.w)));
  float inv_n = fm(-0.5f, n2, 1.5f);
  q = Q4{n.w * inv_n, n.x * inv_n, n.y * inv_n, n.z * inv_n};
  // body-frame Euler equations, gyroscopic term with the old rates
  float nwx = fm(-P.b_w[0], w.y * w.z, w.x + W.ax);
  float nwy = fm(-P.b_w[1], w.z * w.x, w.y + W.ay);
  float nwz = fm(-P.b_w[2], w.x * w.y, w.z + W.az);
  w = V3{nwx, nwy, nwz};
}

__device__ __forceinline__ void substep(const DdqcEnvParams& P, const Wrench& W, V3& p, Q4& q, V3& v, V3& w) {
  substep_pose_rates(P, W, p, q, v, w);
  substep_velocity(P, W, q, v);
}

// `decimation` substeps under one zero-order-held rotor command.  DDQC_ENV_DEFER_VELOCITY = 1 defers the velocity half
// of substep k to the top of substep k + 1 (beside the quaternion chain instead of behind it; bit-identical) - measured
// and left off: 63.96-64.04 us per 2^20-env step against 63.63-63.67 us for the plain loop (the 32 resident warps
// already cover that latency; the rotated loop only lengthens the live ranges).
#ifndef DDQC_ENV_DEFER_VELOCITY
#define DDQC_ENV_DEFER_VELOCITY 0
#endif
__device__ __forceinline__ void substeps(const DdqcEnvParams& P, const Wrench& W, V3& p, Q4& q, V3& v, V3& w) {
#if DDQC_ENV_DEFER_VELOCITY
  for (int sidx = 0; sidx < P.decimation; ++sidx) {
    if (sidx > 0) substep_velocity(P, W, q, v);  // deferred half of the previous substep
    substep_pose_rates(P, W, p, q, v, w);
  }
  if (P.decimation > 0) substep_velocity(P, W, q, v);
#else
  for (int sidx = 0; sidx < P.decimation; ++sidx) substep(P, W, p, q, v, w);
#endif
}

__device__ __forceinline__ float sumsq3(V3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }


}  // namespace ddqc
