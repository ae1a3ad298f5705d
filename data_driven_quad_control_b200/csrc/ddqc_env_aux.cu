// ddqc_env_aux.cu -- rare-path environment kernels: external reset_idx, state save / restore,
// target update, body-frame velocity materialisation, pending PID-reset flush.
// (reference: src/data_driven_quad_control/envs/hover_env.py:556-595, 667-753.)
// Compiled with -fmad=false like ddqc_env.cu (the body-frame transforms must match it).
#include "ddqc_device.cuh"

using namespace ddqc;

namespace {

constexpr int kThreads = 256;

// Shared tail of reset / restore launches: fold the per-copy accumulators into the
// extras["episode"] ring (mean over the touched envs / episode_length_s, hover_env.py:583-589).
__device__ void finalize_episode_stats(DdqcEnvCtl* ctl, float episode_length_s, bool bump_rng, uint32_t ep_row) {
  float sums[DDQC_NUM_REWARDS];
  unsigned int cnt = 0;
  for (int k = 0; k < DDQC_NUM_REWARDS; ++k) sums[k] = 0.0f;
  for (int c = 0; c < DDQC_STAT_COPIES; ++c) {
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
      sums[k] += __ldcg(&ctl->acc_sums[c][k]);
      ctl->acc_sums[c][k] = 0.0f;
    }
    ctl->acc_sums[c][7] = 0.0f;
    cnt += __ldcg(&ctl->acc_reset[c]);
    ctl->acc_reset[c] = 0u;
  }
  const uint32_t row = ep_row % DDQC_EP_RING;  // named by the host (DdqcEnvOut.ep_row)
  const uint32_t prev = ctl->ep_row;
  for (int k = 0; k < DDQC_NUM_REWARDS; ++k)
    ctl->ep_ring[row][k] = cnt ? (sums[k] / (float)cnt) / episode_length_s : ctl->ep_ring[prev][k];
  ctl->ep_ring[row][7] = (float)cnt;
  ctl->ep_row = row;
  if (bump_rng) ctl->rng_event += 1;
  ctl->blocks_done = 0u;
}

__global__ void __launch_bounds__(kThreads)
reset_idx_kernel(const DdqcEnvParams P, const DdqcEnvState S, const DdqcEnvOut O, DdqcEnvCtl* __restrict__ ctl,
                 const int64_t* __restrict__ idx, const int n_idx, const int n, const int A) {
  __shared__ float s_sums[DDQC_NUM_REWARDS];
  __shared__ unsigned int s_cnt;
  __shared__ bool s_last;
  if (threadIdx.x < DDQC_NUM_REWARDS) s_sums[threadIdx.x] = 0.0f;
  if (threadIdx.x == 0) s_cnt = 0u;
  __syncthreads();
  const uint64_t ev = ctl->rng_event;
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t < n_idx) {
    const int e = (int)idx[t];
    if (e >= 0 && e < n) {
      for (int c = 0; c < 3; ++c) {
        S.pos[c * n + e] = P.init_pos[c];
        S.vel[c * n + e] = 0.0f;
        S.ang[c * n + e] = 0.0f;
      }
      for (int c = 0; c < 4; ++c) S.quat[c * n + e] = P.init_quat[c];
      for (int j = 0; j < A; ++j) S.last_actions[j * n + e] = 0.0f;
      S.episode_length[e] = 0;
      S.hover_counter[e] = 0;
      if ((P.flags & DDQC_F_INDEPENDENT_ENVS) && P.action_type != DDQC_ACT_ROTOR_RPMS && S.pid_integral && S.pid_prev_error) {
        for (int c = 0; c < 3; ++c) {  // independent instances: the rate PID of the reset envs only
          S.pid_integral[c * n + e] = 0.0f;
          S.pid_prev_error[c * n + e] = 0.0f;
        }
      }
      if (O.reset) O.reset[e] = 1;
      for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
        atomicAdd(&s_sums[k], S.episode_sums[k * n + e]);
        S.episode_sums[k * n + e] = 0.0f;
      }
      atomicAdd(&s_cnt, 1u);
      V3 c = resample_command(P, (uint32_t)e, ev, DDQC_STREAM_RESET);
      S.commands[e] = c.x;
      S.commands[n + e] = c.y;
      S.commands[2 * n + e] = c.z;
    }
  }
  __syncthreads();
  const int copy = blockIdx.x % DDQC_STAT_COPIES;
  if (threadIdx.x < DDQC_NUM_REWARDS) {
    atomicAdd(&ctl->acc_sums[copy][threadIdx.x], s_sums[threadIdx.x]);
  } else if (threadIdx.x == DDQC_NUM_REWARDS) {
    atomicAdd(&ctl->acc_reset[copy], s_cnt);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctl->blocks_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    finalize_episode_stats(ctl, P.episode_length_s, true, O.ep_row);
    // ctbr_controller.reset(): whole-batch PID zero (hover_env.py:592-593), applied lazily
    if (P.action_type != DDQC_ACT_ROTOR_RPMS && !(P.flags & DDQC_F_INDEPENDENT_ENVS)) ctl->pid_reset_pending = 1u;
  }
}

__global__ void __launch_bounds__(kThreads)
get_state_kernel(const DdqcEnvState S, const int64_t* __restrict__ idx, const int n_idx, const int n, const int A,
                 float* __restrict__ base_pos, float* __restrict__ base_quat, float* __restrict__ base_lin_vel,
                 float* __restrict__ base_ang_vel, float* __restrict__ commands, int32_t* __restrict__ episode_length,
                 float* __restrict__ last_actions, const float* __restrict__ ang_override,
                 const float* __restrict__ lin_override) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_idx) return;
  const int e = (int)idx[t];
  if (e < 0 || e >= n) return;
  V3 p{S.pos[e], S.pos[n + e], S.pos[2 * n + e]};
  Q4 q{S.quat[e], S.quat[n + e], S.quat[2 * n + e], S.quat[3 * n + e]};
  V3 v{S.vel[e], S.vel[n + e], S.vel[2 * n + e]};
  V3 w{S.ang[e], S.ang[n + e], S.ang[2 * n + e]};
  Q4 qc{q.w, -q.x, -q.y, -q.z};
  V3 lin_b = lin_override ? V3{lin_override[e], lin_override[n + e], lin_override[2 * n + e]} : rotate_by_quat(v, qc);
  V3 ang_b = ang_override ? V3{ang_override[e], ang_override[n + e], ang_override[2 * n + e]}
                          : rotate_by_quat(rotate_by_quat(w, q), qc);
  base_pos[t * 3 + 0] = p.x;
  base_pos[t * 3 + 1] = p.y;
  base_pos[t * 3 + 2] = p.z;
  base_quat[t * 4 + 0] = q.w;
  base_quat[t * 4 + 1] = q.x;
  base_quat[t * 4 + 2] = q.y;
  base_quat[t * 4 + 3] = q.z;
  base_lin_vel[t * 3 + 0] = lin_b.x;
  base_lin_vel[t * 3 + 1] = lin_b.y;
  base_lin_vel[t * 3 + 2] = lin_b.z;
  base_ang_vel[t * 3 + 0] = ang_b.x;
  base_ang_vel[t * 3 + 1] = ang_b.y;
  base_ang_vel[t * 3 + 2] = ang_b.z;
  for (int c = 0; c < 3; ++c) commands[t * 3 + c] = S.commands[c * n + e];
  episode_length[t] = S.episode_length[e];
  for (int j = 0; j < A; ++j) last_actions[t * A + j] = S.last_actions[j * n + e];
}

__global__ void __launch_bounds__(kThreads)
restore_state_kernel(const DdqcEnvParams P, const DdqcEnvState S, const DdqcEnvOut O, DdqcEnvCtl* __restrict__ ctl,
                     const int64_t* __restrict__ idx, const int n_idx, const int n, const int A,
                     const float* __restrict__ base_pos, const float* __restrict__ base_quat,
                     const float* __restrict__ base_lin_vel, const float* __restrict__ base_ang_vel,
                     const float* __restrict__ commands, const int32_t* __restrict__ episode_length,
                     const float* __restrict__ last_actions) {
  __shared__ float s_sums[DDQC_NUM_REWARDS];
  __shared__ unsigned int s_cnt;
  __shared__ bool s_last;
  if (threadIdx.x < DDQC_NUM_REWARDS) s_sums[threadIdx.x] = 0.0f;
  if (threadIdx.x == 0) s_cnt = 0u;
  __syncthreads();
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t < n_idx) {
    const int e = (int)idx[t];
    if (e >= 0 && e < n) {
      for (int c = 0; c < 3; ++c) {
        const float pc = base_pos[t * 3 + c], cm = commands[t * 3 + c];
        S.pos[c * n + e] = pc;
        // the saved commands only enter rel_pos / last_rel_pos: the reference's restore_from_state never writes
        // self.commands (hover_env.py:712-716), so an env whose target changed since the save keeps its CURRENT target
        // body-frame buffers go into the simulator dofs verbatim (hover_env.py:729-735)
        S.vel[c * n + e] = base_lin_vel[t * 3 + c];
        S.ang[c * n + e] = base_ang_vel[t * 3 + c];
        if (O.rel_pos) O.rel_pos[c * n + e] = cm - pc;
        if (O.last_rel_pos) O.last_rel_pos[c * n + e] = cm - pc;
        if (O.last_base_pos) O.last_base_pos[c * n + e] = pc;
        if (O.base_lin_vel) O.base_lin_vel[c * n + e] = base_lin_vel[t * 3 + c];
        if (O.base_ang_vel) O.base_ang_vel[c * n + e] = base_ang_vel[t * 3 + c];
      }
      for (int c = 0; c < 4; ++c) S.quat[c * n + e] = base_quat[t * 4 + c];
      for (int j = 0; j < A; ++j) S.last_actions[j * n + e] = last_actions[t * A + j];
      S.episode_length[e] = episode_length[t];
      if (O.reset) O.reset[e] = 1;
      for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
        atomicAdd(&s_sums[k], S.episode_sums[k * n + e]);
        S.episode_sums[k * n + e] = 0.0f;
      }
      atomicAdd(&s_cnt, 1u);
    }
  }
  __syncthreads();
  const int copy = blockIdx.x % DDQC_STAT_COPIES;
  if (threadIdx.x < DDQC_NUM_REWARDS) {
    atomicAdd(&ctl->acc_sums[copy][threadIdx.x], s_sums[threadIdx.x]);
  } else if (threadIdx.x == DDQC_NUM_REWARDS) {
    atomicAdd(&ctl->acc_reset[copy], s_cnt);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctl->blocks_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    finalize_episode_stats(ctl, P.episode_length_s, false, O.ep_row);
    ctl->pid_reset_pending = 0u;  // load_state() of the whole-batch PID follows on the host side
  }
}

__global__ void __launch_bounds__(kThreads)
update_target_kernel(const DdqcEnvState S, const int64_t* __restrict__ idx, const int n_idx, const int n,
                     const float* __restrict__ target, const int broadcast) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= n_idx) return;
  const int e = (int)idx[t];
  if (e < 0 || e >= n) return;
  const float* src = broadcast ? target : target + (size_t)t * 3;
  S.commands[e] = src[0];
  S.commands[n + e] = src[1];
  S.commands[2 * n + e] = src[2];
}

__global__ void __launch_bounds__(kThreads)
body_vel_kernel(const DdqcEnvState S, const int n, float* __restrict__ lin, float* __restrict__ ang) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  Q4 q{S.quat[i], S.quat[n + i], S.quat[2 * n + i], S.quat[3 * n + i]};
  V3 v{S.vel[i], S.vel[n + i], S.vel[2 * n + i]};
  V3 w{S.ang[i], S.ang[n + i], S.ang[2 * n + i]};
  Q4 qc{q.w, -q.x, -q.y, -q.z};
  V3 lb = rotate_by_quat(v, qc);
  V3 ab = rotate_by_quat(rotate_by_quat(w, q), qc);
  if (lin) {
    lin[i] = lb.x;
    lin[n + i] = lb.y;
    lin[2 * n + i] = lb.z;
  }
  if (ang) {
    ang[i] = ab.x;
    ang[n + i] = ab.y;
    ang[2 * n + i] = ab.z;
  }
}

__global__ void __launch_bounds__(kThreads)
flush_pid_kernel(const DdqcEnvState S, const DdqcEnvCtl* __restrict__ ctl, const int n) {
  if (ctl->pid_reset_pending == 0u) return;
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < 3; ++c) {
    S.pid_integral[c * n + i] = 0.0f;
    S.pid_prev_error[c * n + i] = 0.0f;
  }
}
__global__ void clear_pid_flag_kernel(DdqcEnvCtl* ctl) { ctl->pid_reset_pending = 0u; }

inline int num_actions(int action_type) { return action_type == DDQC_ACT_CTBR_FIXED_YAW ? 3 : 4; }
inline dim3 grid_for(int64_t n) { return dim3((unsigned)((n + kThreads - 1) / kThreads)); }

}  // namespace

extern "C" int ddqc_env_reset_idx(const DdqcEnvParams* P, const DdqcEnvState* S, const DdqcEnvOut* O, DdqcEnvCtl* ctl,
                                  const int64_t* idx, int64_t n_idx, int64_t n_env, void* stream) {
  if (!P || !S || !O || !ctl || !idx) return DDQC_E_NULL;
  if (n_env <= 0 || n_idx < 0 || n_env > 0x7fffffff / 8) return DDQC_E_SHAPE;
  if (n_idx == 0) return 0;  // hover_env.py:557-558
  reset_idx_kernel<<<grid_for(n_idx), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      *P, *S, *O, ctl, idx, (int)n_idx, (int)n_env, num_actions(P->action_type));
  return (int)cudaGetLastError();
}

extern "C" int ddqc_env_get_state(const DdqcEnvParams* P, const DdqcEnvState* S, const int64_t* idx, int64_t n_idx,
                                  int64_t n_env, float* base_pos, float* base_quat, float* base_lin_vel,
                                  float* base_ang_vel, float* commands, int32_t* episode_length, float* last_actions,
                                  const float* ang_override, const float* lin_override, void* stream) {
  if (!P || !S || !idx || !base_pos || !base_quat || !base_lin_vel || !base_ang_vel || !commands || !episode_length ||
      !last_actions)
    return DDQC_E_NULL;
  if (n_env <= 0 || n_idx < 0) return DDQC_E_SHAPE;
  if (n_idx == 0) return 0;
  get_state_kernel<<<grid_for(n_idx), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      *S, idx, (int)n_idx, (int)n_env, num_actions(P->action_type), base_pos, base_quat, base_lin_vel, base_ang_vel,
      commands, episode_length, last_actions, ang_override, lin_override);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_env_restore_state(const DdqcEnvParams* P, const DdqcEnvState* S, const DdqcEnvOut* O,
                                      DdqcEnvCtl* ctl, const int64_t* idx, int64_t n_idx, int64_t n_env,
                                      const float* base_pos, const float* base_quat, const float* base_lin_vel,
                                      const float* base_ang_vel, const float* commands, const int32_t* episode_length,
                                      const float* last_actions, void* stream) {
  if (!P || !S || !O || !ctl || !idx || !base_pos || !base_quat || !base_lin_vel || !base_ang_vel || !commands ||
      !episode_length || !last_actions)
    return DDQC_E_NULL;
  if (n_env <= 0 || n_idx < 0) return DDQC_E_SHAPE;
  if (n_idx == 0) return 0;
  restore_state_kernel<<<grid_for(n_idx), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      *P, *S, *O, ctl, idx, (int)n_idx, (int)n_env, num_actions(P->action_type), base_pos, base_quat, base_lin_vel,
      base_ang_vel, commands, episode_length, last_actions);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_env_update_target(const DdqcEnvState* S, const int64_t* idx, int64_t n_idx, int64_t n_env,
                                      const float* target_pos, int broadcast, void* stream) {
  if (!S || !idx || !target_pos) return DDQC_E_NULL;
  if (n_env <= 0 || n_idx < 0) return DDQC_E_SHAPE;
  if (n_idx == 0) return 0;
  update_target_kernel<<<grid_for(n_idx), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      *S, idx, (int)n_idx, (int)n_env, target_pos, broadcast);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_env_body_vel(const DdqcEnvState* S, int64_t n_env, float* base_lin_vel, float* base_ang_vel,
                                 void* stream) {
  if (!S) return DDQC_E_NULL;
  if (n_env <= 0) return DDQC_E_SHAPE;
  body_vel_kernel<<<grid_for(n_env), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*S, (int)n_env, base_lin_vel,
                                                                                       base_ang_vel);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_env_flush_pid_reset(const DdqcEnvState* S, DdqcEnvCtl* ctl, int64_t n_env, void* stream) {
  if (!S || !ctl) return DDQC_E_NULL;
  if (!S->pid_integral || !S->pid_prev_error) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  flush_pid_kernel<<<grid_for(n_env), kThreads, 0, st>>>(*S, ctl, (int)n_env);
  clear_pid_flag_kernel<<<1, 1, 0, st>>>(ctl);
  return (int)cudaGetLastError();
}

extern "C" int ddqc_abi_version(void) { return DDQC_ABI_VERSION; }

extern "C" int64_t ddqc_struct_size(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(DdqcEnvParams);
    case 1: return (int64_t)sizeof(DdqcEnvCtl);
    case 2: return (int64_t)sizeof(DdqcEnvState);
    case 3: return (int64_t)sizeof(DdqcEnvOut);
    case 4: return (int64_t)sizeof(DdqcMpcConfig);
    case 5: return (int64_t)sizeof(DdqcMpcAux);
    default: return -1;
  }
}
