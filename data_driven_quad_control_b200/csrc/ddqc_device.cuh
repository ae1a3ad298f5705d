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

__device__ __forceinline__ Q4 quat_mul(Q4 a, Q4 b) {
  Q4 r;
  r.w = ((a.w * b.w - a.x * b.x) - a.y * b.y) - a.z * b.z;
  r.x = ((a.w * b.x + a.x * b.w) + a.y * b.z) - a.z * b.y;
  r.y = ((a.w * b.y - a.x * b.z) + a.y * b.w) + a.z * b.x;
  r.z = ((a.w * b.z + a.x * b.y) - a.y * b.x) + a.z * b.w;
  return r;
}

// genesis.utils.geom.transform_by_quat: v + w t + qv x t, t = 2 qv x v
__device__ __forceinline__ V3 rotate_by_quat(V3 v, Q4 q) {
  float tx = (q.y * v.z - q.z * v.y) * 2.0f;
  float ty = (q.z * v.x - q.x * v.z) * 2.0f;
  float tz = (q.x * v.y - q.y * v.x) * 2.0f;
  V3 r;
  r.x = (v.x + q.w * tx) + (q.y * tz - q.z * ty);
  r.y = (v.y + q.w * ty) + (q.z * tx - q.x * tz);
  r.z = (v.z + q.w * tz) + (q.x * ty - q.y * tx);
  return r;
}

__device__ __forceinline__ float horner7(const float* c, float s) {
  float acc = s * c[6] + c[5];
  acc = acc * s + c[4];
  acc = acc * s + c[3];
  acc = acc * s + c[2];
  acc = acc * s + c[1];
  acc = acc * s + c[0];
  return acc;
}

__device__ __forceinline__ float clipf(float x, float lo, float hi) {
  return x < lo ? lo : (x > hi ? hi : x);  // NaN passes through, like torch.clip
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

// ---- one rigid-body substep (oracle/drone_physics.py: substep) ------------------------
__device__ __forceinline__ void substep(const DdqcEnvParams& P, V3& p, Q4& q, V3& v, V3& w, const float rpm[4]) {
  float t0 = (rpm[0] * rpm[0]) * P.kf, t1 = (rpm[1] * rpm[1]) * P.kf;
  float t2 = (rpm[2] * rpm[2]) * P.kf, t3 = (rpm[3] * rpm[3]) * P.kf;
  float fz = ((t0 + t1) + t2) + t3;
  float tau_x = ((t0 * P.rotor_y[0] + t1 * P.rotor_y[1]) + t2 * P.rotor_y[2]) + t3 * P.rotor_y[3];
  float tau_y = ((t0 * (-P.rotor_x[0]) + t1 * (-P.rotor_x[1])) + t2 * (-P.rotor_x[2])) + t3 * (-P.rotor_x[3]);
  float m0 = ((rpm[0] * rpm[0]) * P.km) * P.spin[0], m1 = ((rpm[1] * rpm[1]) * P.km) * P.spin[1];
  float m2 = ((rpm[2] * rpm[2]) * P.km) * P.spin[2], m3 = ((rpm[3] * rpm[3]) * P.km) * P.spin[3];
  float tau_z = ((m0 + m1) + m2) + m3;

  float jwx = w.x * P.jxx, jwy = w.y * P.jyy, jwz = w.z * P.jzz;
  float gx = w.y * jwz - w.z * jwy;
  float gy = w.z * jwx - w.x * jwz;
  float gz = w.x * jwy - w.y * jwx;
  w.x = w.x + ((tau_x - gx) * P.inv_jxx) * P.dt;
  w.y = w.y + ((tau_y - gy) * P.inv_jyy) * P.dt;
  w.z = w.z + ((tau_z - gz) * P.inv_jzz) * P.dt;

  float r13 = (q.x * q.z + q.w * q.y) * 2.0f;
  float r23 = (q.y * q.z - q.w * q.x) * 2.0f;
  float r33 = ((q.w * q.w - q.x * q.x) - q.y * q.y) + q.z * q.z;
  float acc = fz * P.inv_mass;
  v.x = v.x + (r13 * acc) * P.dt;
  v.y = v.y + (r23 * acc) * P.dt;
  v.z = v.z + (r33 * acc - P.gravity) * P.dt;

  p.x = p.x + v.x * P.dt;
  p.y = p.y + v.y * P.dt;
  p.z = p.z + v.z * P.dt;

  float hx = w.x * P.half_dt, hy = w.y * P.half_dt, hz = w.z * P.half_dt;
  float s = (hx * hx + hy * hy) + hz * hz;
  float c = horner7(P.cos_coef, s);
  float k = horner7(P.sinc_coef, s);
  Q4 n = quat_mul(q, Q4{c, k * hx, k * hy, k * hz});
  float inv_n = 1.0f / sqrtf(((n.w * n.w + n.x * n.x) + n.y * n.y) + n.z * n.z);
  q = Q4{n.w * inv_n, n.x * inv_n, n.y * inv_n, n.z * inv_n};
}

__device__ __forceinline__ float sumsq3(V3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }


}  // namespace ddqc
