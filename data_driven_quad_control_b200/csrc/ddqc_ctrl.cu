// ddqc_ctrl.cu -- stand-alone vectorized controllers for sm_100a:
//   VectorizedPIDController.compute   (utilities/vectorized_pid_controller.py:97-126)
//   DroneCTBRController.compute       (controllers/ctbr/ctbr_controller.py:221-264)
//   DroneTrackingController.compute   (controllers/tracking/tracking_controller.py:90-164)
// One thread per drone; inputs are addressed with explicit element strides so (N,k) tensors and
// their transposed struct-of-arrays views are both read without a copy.  Compiled with
// -fmad=false: the op order follows the reference's elementwise torch expressions; tensor / scalar
// divisions follow torch's CUDA semantics, x * (1.0f / s).
#include "ddqc_device.cuh"

using namespace ddqc;

namespace {

constexpr int kThreads = 256;

struct Gains {
  float kp[8], ki[8], kd[8];
};
struct Mixer {
  float m[16];
};

__device__ __forceinline__ float pid_update(float sp, float meas, float& integ, float& prev, float kp, float ki,
                                            float kd, float dt, float inv_dt) {
  float err = sp - meas;
  integ = integ + err * dt;
  float deriv = (err - prev) * inv_dt;
  prev = err;
  return (kp * err + ki * integ) + kd * deriv;
}

__global__ void __launch_bounds__(kThreads)
pid_kernel(const float* __restrict__ sp, int64_t sp_rs, int64_t sp_cs, const float* __restrict__ ms, int64_t ms_rs,
           int64_t ms_cs, float* __restrict__ integral, float* __restrict__ prev_error, const Gains g, const float dt,
           float* __restrict__ out, const int64_t n, const int k) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  for (int j = 0; j < k; ++j) {
    float integ = integral[i * k + j], prev = prev_error[i * k + j];
    out[i * k + j] = pid_update(sp[i * sp_rs + j * sp_cs], ms[i * ms_rs + j * ms_cs], integ, prev, g.kp[j], g.ki[j],
                                g.kd[j], dt, 1.0f / dt);
    integral[i * k + j] = integ;
    prev_error[i * k + j] = prev;
  }
}

__global__ void __launch_bounds__(kThreads)
ctbr_kernel(const float* __restrict__ meas, int64_t m_rs, int64_t m_cs, const float* __restrict__ sp, int64_t s_rs,
            int64_t s_cs, const float* __restrict__ thrust, int64_t t_s, float* __restrict__ integral,
            float* __restrict__ prev_error, const Mixer M, const Gains g, const float dt, const float inv_sqrt_kf,
            float* __restrict__ rpm, const int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float c[4];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float integ = integral[i * 3 + j], prev = prev_error[i * 3 + j];
    c[j] = pid_update(sp[i * s_rs + j * s_cs], meas[i * m_rs + j * m_cs], integ, prev, g.kp[j], g.ki[j], g.kd[j], dt, 1.0f / dt);
    integral[i * 3 + j] = integ;
    prev_error[i * 3 + j] = prev;
  }
  c[3] = thrust[i * t_s];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float f = ((M.m[4 * r + 0] * c[0] + M.m[4 * r + 1] * c[1]) + M.m[4 * r + 2] * c[2]) + M.m[4 * r + 3] * c[3];
    f = f < 0.0f ? 0.0f : f;
    rpm[i * 4 + r] = sqrt_nonneg(f) * inv_sqrt_kf;
  }
}

__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// torch.nn.functional.normalize(v, dim=1): v / max(||v||_2, 1e-12)
__device__ __forceinline__ V3 normalize3(V3 a) {
  float nrm = sqrtf((a.x * a.x + a.y * a.y) + a.z * a.z);
  nrm = nrm < 1e-12f ? 1e-12f : nrm;
  return V3{a.x / nrm, a.y / nrm, a.z / nrm};
}

__global__ void __launch_bounds__(kThreads)
tracking_kernel(const float* __restrict__ pos, int64_t p_rs, int64_t p_cs, const float* __restrict__ quat,
                int64_t q_rs, int64_t q_cs, const float* __restrict__ pos_sp, int64_t ps_rs, int64_t ps_cs,
                const float* __restrict__ quat_sp, int64_t qs_rs, int64_t qs_cs, float* __restrict__ integral,
                float* __restrict__ prev_error, const Gains g, const float dt, const float mass, const float kR,
                float* __restrict__ out, const int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  float acc[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float integ = integral[i * 3 + j], prev = prev_error[i * 3 + j];
    acc[j] = pid_update(pos_sp[i * ps_rs + j * ps_cs], pos[i * p_rs + j * p_cs], integ, prev, g.kp[j], g.ki[j],
                        g.kd[j], dt, 1.0f / dt);
    integral[i * 3 + j] = integ;
    prev_error[i * 3 + j] = prev;
  }
  acc[2] = acc[2] + 9.81f;  // gravity compensation (tracking_controller.py:72-76,118)
  V3 f{acc[0] * mass, acc[1] * mass, acc[2] * mass};

  // R(q) of the measured attitude (math_utils.py:23-61)
  const float w = quat[i * q_rs], x = quat[i * q_rs + q_cs], y = quat[i * q_rs + 2 * q_cs], z = quat[i * q_rs + 3 * q_cs];
  const float xx = x * x, yy = y * y, zz = z * z, ww = w * w;
  const float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  float R[3][3];
  R[0][0] = ((ww + xx) - yy) - zz;
  R[0][1] = 2.0f * (xy - wz);
  R[0][2] = 2.0f * (xz + wy);
  R[1][0] = 2.0f * (xy + wz);
  R[1][1] = ((ww - xx) + yy) - zz;
  R[1][2] = 2.0f * (yz - wx);
  R[2][0] = 2.0f * (xz - wy);
  R[2][1] = 2.0f * (yz + wx);
  R[2][2] = ((ww - xx) - yy) + zz;
  // thrust = (R f)[2]   -- R, not R^T, as in the reference (tracking_controller.py:124-126)
  const float thrust = (R[2][0] * f.x + R[2][1] * f.y) + R[2][2] * f.z;

  V3 z_des = normalize3(f);
  const float sw = quat_sp[i * qs_rs], sx = quat_sp[i * qs_rs + qs_cs], sy = quat_sp[i * qs_rs + 2 * qs_cs],
              sz = quat_sp[i * qs_rs + 3 * qs_cs];
  const float yaw_des = atan2f(2.0f * (sw * sz + sx * sy), 1.0f - 2.0f * (sy * sy + sz * sz));
  V3 x_yaw{cosf(yaw_des), sinf(yaw_des), 0.0f};
  V3 y_des = normalize3(cross3(z_des, x_yaw));
  V3 x_des = normalize3(cross3(y_des, z_des));
  float Rd[3][3] = {{x_des.x, y_des.x, z_des.x}, {x_des.y, y_des.y, z_des.y}, {x_des.z, y_des.z, z_des.z}};
  // M = Rd^T R ; e_R = vee(0.5 (M - M^T))
  float M[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) M[a][b] = (Rd[0][a] * R[0][b] + Rd[1][a] * R[1][b]) + Rd[2][a] * R[2][b];
  const float e0 = 0.5f * (M[2][1] - M[1][2]);
  const float e1 = 0.5f * (M[0][2] - M[2][0]);
  const float e2 = 0.5f * (M[1][0] - M[0][1]);
  out[i * 4 + 0] = thrust;
  out[i * 4 + 1] = -kR * e0;
  out[i * 4 + 2] = -kR * e1;
  out[i * 4 + 3] = -kR * e2;
}

inline dim3 grid_for(int64_t n) { return dim3((unsigned)((n + kThreads - 1) / kThreads)); }
inline Gains load_gains(const float* g, int k) {
  Gains out{};
  for (int j = 0; j < k; ++j) {
    out.kp[j] = g[3 * j + 0];
    out.ki[j] = g[3 * j + 1];
    out.kd[j] = g[3 * j + 2];
  }
  return out;
}

}  // namespace

extern "C" int ddqc_pid_compute(const float* sp, int64_t sp_rs, int64_t sp_cs, const float* ms, int64_t ms_rs,
                                int64_t ms_cs, float* integral, float* prev_error, const float* gains_host, float dt,
                                float* out, int64_t n, int32_t k, void* stream) {
  if (!sp || !ms || !integral || !prev_error || !gains_host || !out) return DDQC_E_NULL;
  if (n < 0 || k < 1 || k > 8) return DDQC_E_SHAPE;
  if (n == 0) return 0;
  pid_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      sp, sp_rs, sp_cs, ms, ms_rs, ms_cs, integral, prev_error, load_gains(gains_host, k), dt, out, n, k);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_ctbr_compute(const float* meas, int64_t m_rs, int64_t m_cs, const float* sp, int64_t s_rs,
                                 int64_t s_cs, const float* thrust, int64_t t_s, float* integral, float* prev_error,
                                 const float* mixer_host, const float* gains_host, float dt, float inv_sqrt_kf,
                                 float* rpm, int64_t n, void* stream) {
  if (!meas || !sp || !thrust || !integral || !prev_error || !mixer_host || !gains_host || !rpm) return DDQC_E_NULL;
  if (n < 0) return DDQC_E_SHAPE;
  if (n == 0) return 0;
  Mixer M;
  for (int j = 0; j < 16; ++j) M.m[j] = mixer_host[j];
  ctbr_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      meas, m_rs, m_cs, sp, s_rs, s_cs, thrust, t_s, integral, prev_error, M, load_gains(gains_host, 3), dt,
      inv_sqrt_kf, rpm, n);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_tracking_compute(const float* pos, int64_t p_rs, int64_t p_cs, const float* quat, int64_t q_rs,
                                     int64_t q_cs, const float* pos_sp, int64_t ps_rs, int64_t ps_cs,
                                     const float* quat_sp, int64_t qs_rs, int64_t qs_cs, float* integral,
                                     float* prev_error, const float* gains_host, float dt, float mass, float kR,
                                     float* out, int64_t n, void* stream) {
  if (!pos || !quat || !pos_sp || !quat_sp || !integral || !prev_error || !gains_host || !out) return DDQC_E_NULL;
  if (n < 0) return DDQC_E_SHAPE;
  if (n == 0) return 0;
  tracking_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pos, p_rs, p_cs, quat, q_rs, q_cs, pos_sp, ps_rs, ps_cs, quat_sp, qs_rs, qs_cs, integral, prev_error,
      load_gains(gains_host, 3), dt, mass, kR, out, n);
  return (int)cudaGetLastError();
}
