// ddqc_env.cu -- fused vectorized drone-environment step for sm_100a.
//
// One thread owns one environment: it loads the env's struct-of-arrays state with fully
// coalesced 32-bit loads (one 128 B line per warp per component), keeps it in registers
// across the action decode, the CTBR rate controller, the `decimation` rigid-body substeps,
// the termination / auto-reset logic, the rewards and the observation, and streams the new
// state and the outputs back.  The kernel is HBM-bound (357 B / env-step of algorithmic
// traffic for CTBR_FIXED_YAW, ~1 kFLOP), so nothing here goes near the tensor cores.
//
// Arithmetic contract: this translation unit is compiled with -fmad=false and mirrors,
// operation for operation (explicit fm() = fused, everything else separately rounded), the fp32
// restatement in oracle/env_oracle.py +
// oracle/drone_physics.py, so positions, rewards and reset masks are bit-identical to
// the oracle (libm calls -- atan2/asin/exp/log/cos -- are within 2 ulp).
//
// Reference behaviour reproduced (src/data_driven_quad_control/envs/hover_env.py:376-547,
// 556-595 and SURVEY.md Appendix A1-A24), including the two batch-global quirks:
//   A17  any reset in step t zeroes every env's rate-PID state before step t+1
//        -> `ctl->pid_reset_pending`, written by the last CTA of step t, read in t+1;
//   A16  rel_pos / last_rel_pos are refreshed for all envs only on steps with >=1 reset
//        -> envs whose target was resampled without a reset append the "fresh" variant of
//           their reward / sums / obs to a fix-up list that the last CTA applies iff
//           n_reset > 0.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ddqc_device.cuh"

using namespace ddqc;

namespace {

// threads per CTA of the two launch shapes (1024 threads = 32 warps of 64 registers per SM either way)
#ifndef DDQC_ENV_THREADS_FUSED
#define DDQC_ENV_THREADS_FUSED 128
#endif
#ifndef DDQC_ENV_THREADS_SPLIT
#define DDQC_ENV_THREADS_SPLIT 256
#endif
// Measured and dropped: prefetch.global.L2 of the state of the env one wave (or half / two waves) of resident CTAs
// ahead - 0.0734 / 0.0745 / 0.0803 ms per 2^20 envs against 0.0735 ms without (eager launches): the loads already
// cover the DRAM latency with 32 warps per SM, the ~40 extra instructions per thread only cost issue slots.
// Where a step's statistics are folded (template parameter FUSED of the step kernel):
//   FUSED   the last CTA of the step kernel does it (bar.sync + device fence + ticket per CTA): one launch per step,
//           the lowest latency for small batches.  For large batches ncu showed warp 0 of every CTA waiting on the
//           fence for half as long again as the CTA's work takes, with the registers of the three exited warps still
//           held (CTA residency 1.5x the work), plus 9 % of the warp samples at the barrier: 70.8 us per 2^20 envs;
//   !FUSED  a second one-warp launch (env_finalize_kernel) folds them; the step kernel has no barrier, no fence and no
//           ticket.  DDQC_ENV_FINALIZE selects how the two are chained: 1 = stream order (66.0 us), 2 = programmatic
//           dependent launch (64.7 us; 64.0 us with 256-thread CTAs): the finalisation is scheduled while the step
//           kernel drains and the next step kernel while the finalisation runs; each waits (griddepcontrol.wait)
//           before touching memory.
// ddqc_env_step picks FUSED for n_env <= kFusedMaxEnvs (override: environment variable DDQC_ENV_FUSED_MAX_N).
#ifndef DDQC_ENV_FINALIZE
#define DDQC_ENV_FINALIZE 2
#endif
#ifndef DDQC_ENV_FUSED_MAX_N
#define DDQC_ENV_FUSED_MAX_N 32768
#endif
// State loads bypass L1 (ld.global.cg): every datum is read once, and the loads issued before griddepcontrol.wait must
// not be served from lines an earlier grid left in this SM's L1.  Measured and dropped: streaming (evict-first) hints on
// the outputs, on the state stores and on the state loads - 64.2 / 64.3 / 64.6 us per 2^20 envs against 63.9 us without.
template <typename T>
__device__ __forceinline__ T ld_state(const T* p) {
  return __ldcg(p);
}
template <typename T>
__device__ __forceinline__ void st_state(T* p, T v) {
  *p = v;
}
template <typename T>
__device__ __forceinline__ void st_out(T* p, T v) {
  *p = v;
}
constexpr int kThreadsFused = DDQC_ENV_THREADS_FUSED;
constexpr int kThreadsSplit = DDQC_ENV_THREADS_SPLIT;

// Programmatic dependent launch (sm_90+): no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() {
#if DDQC_ENV_FINALIZE == 2
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if DDQC_ENV_FINALIZE == 2
  asm volatile("griddepcontrol.launch_dependents;");
#endif
}

// ---- step finalisation ------------------------------------------------------------------
// One warp, after every env of the step has been processed: fold the per-copy accumulators
// (lane k < 8: sums, 8: resets, 9: time-outs), apply the A15/A16 fix-up list iff the step saw a
// reset, write the extras["episode"] ring row named by the host (O.ep_row), raise the A17 flag,
// bump the Philox event counter.
template <int NOBS, bool LEAN, bool USES_CTBR>
__device__ __forceinline__ void finalize_step(const DdqcEnvParams& P, const DdqcEnvState& S, const DdqcEnvOut& O,
                                              DdqcEnvCtl* __restrict__ ctl, const int n, const int lane) {
  const float obs_noise_std = LEAN ? 0.0f : P.obs_noise_std;
  const bool indep = (P.flags & DDQC_F_INDEPENDENT_ENVS) != 0u;
  const uint64_t ev = ctl->rng_event;
  // lane c folds copy c of the accumulators (two 16-byte loads + two counters, all 32 copies in flight at once) and
  // clears it; a butterfly leaves every total in every lane
  static_assert(DDQC_STAT_COPIES == 32, "one accumulator copy per lane");
  float f_sums[8];
  unsigned int f_cnt[3];
  {
    float4* acc = reinterpret_cast<float4*>(&ctl->acc_sums[lane][0]);
    const float4 lo = __ldcg(acc), hi = __ldcg(acc + 1);
    f_cnt[0] = __ldcg(&ctl->acc_reset[lane]);
    f_cnt[1] = __ldcg(&ctl->acc_timeout[lane]);
    f_cnt[2] = __ldcg(&ctl->acc_ground[lane]);
    acc[0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    acc[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    ctl->acc_reset[lane] = 0u;
    ctl->acc_timeout[lane] = 0u;
    ctl->acc_ground[lane] = 0u;
    f_sums[0] = lo.x, f_sums[1] = lo.y, f_sums[2] = lo.z, f_sums[3] = lo.w;
    f_sums[4] = hi.x, f_sums[5] = hi.y, f_sums[6] = hi.z, f_sums[7] = hi.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f_sums[k] += __shfl_xor_sync(0xffffffffu, f_sums[k], o);
      f_cnt[0] += __shfl_xor_sync(0xffffffffu, f_cnt[0], o);
      f_cnt[1] += __shfl_xor_sync(0xffffffffu, f_cnt[1], o);
      f_cnt[2] += __shfl_xor_sync(0xffffffffu, f_cnt[2], o);
    }
  }
  const unsigned int n_reset = f_cnt[0];
  const unsigned int n_fix = __ldcg(&ctl->fixup_count);

  // A16: some env reset this step -> envs that only resampled their target see fresh rel_pos
  if (n_reset > 0 && !indep) {
    for (unsigned int s = lane; s < n_fix; s += 32) {
      int e = __ldcg(&O.fixup_idx[s]);
      const float* fv = O.fixup_val + (size_t)s * 8;
      O.rew[e] = __ldcg(&fv[0]);
      S.episode_sums[0 * n + e] = __ldcg(&fv[1]);
      S.episode_sums[1 * n + e] = __ldcg(&fv[2]);
      float o0 = __ldcg(&fv[3]), o1 = __ldcg(&fv[4]), o2 = __ldcg(&fv[5]);
      if (obs_noise_std > 0.0f) {  // same noise words as the main pass
        float z[13];
        philox_normals<13>((uint32_t)e, ev, DDQC_STREAM_OBS_NOISE, P.seed, z);
        o0 = o0 + z[0] * obs_noise_std;
        o1 = o1 + z[1] * obs_noise_std;
        o2 = o2 + z[2] * obs_noise_std;
      }
      O.obs[(size_t)e * NOBS + 0] = o0;
      O.obs[(size_t)e * NOBS + 1] = o1;
      O.obs[(size_t)e * NOBS + 2] = o2;
      if (!LEAN && O.rel_pos && O.last_rel_pos && O.last_base_pos) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float cm = __ldcg(&S.commands[c * n + e]);
          O.rel_pos[c * n + e] = cm - __ldcg(&S.pos[c * n + e]);
          O.last_rel_pos[c * n + e] = cm - __ldcg(&O.last_base_pos[c * n + e]);
        }
      }
    }
  }

  __syncwarp();
  const uint32_t row = O.ep_row % DDQC_EP_RING;  // named by the host, so its extras["episode"] views cannot drift
  const uint32_t prev_row = ctl->ep_row;
  if (lane < DDQC_NUM_REWARDS) {  // lane k: reward term k of the ring row and of the running totals
    float val = ctl->ep_ring[prev_row][lane];
    float mine = 0.0f;
#pragma unroll
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) mine = (lane == k) ? f_sums[k] : mine;
    if (n_reset > 0) {
      val = (mine / (float)n_reset) / P.episode_length_s;
      ctl->total_episode_sums[lane] += (double)mine;
    }
    ctl->ep_ring[row][lane] = val;
  } else if (lane == 7) {
    ctl->ep_ring[row][7] = (float)n_reset;
    ctl->total_reward += (double)f_sums[7];
  } else if (lane == 8) {
    ctl->total_resets += n_reset;
    ctl->total_timeouts += f_cnt[1];
    ctl->total_env_steps += (uint64_t)n;
    ctl->total_below_ground += f_cnt[2];
  } else if (lane == 9) {
    ctl->n_reset_last = n_reset;
    ctl->pid_reset_pending = (USES_CTBR && !indep) ? (n_reset > 0 ? 1u : 0u) : 0u;
    ctl->step_count += 1;
    ctl->rng_event = ev + 1;
    ctl->fixup_count = 0u;
    ctl->blocks_done = 0u;
  }
  __syncwarp();  // every lane has read ctl->ep_row
  if (lane == 10) ctl->ep_row = row;
}

// ---- the fused step kernel -------------------------------------------------------------
// LEAN = the throughput configuration: no Euler angles needed (thresholds >= 180 deg, yaw reward
// scale 0), no actuator / observation noise, no materialised reference buffers, no measurement
// override.  It is the same code with those branches compiled out (fewer registers, no spills).
template <int ACT, bool LEAN, bool FUSED, int THREADS>
// (the full-featured instantiation gets 80 registers instead of 64: it spilled ~250 bytes per thread at 64)
__global__ void __launch_bounds__(THREADS, (LEAN ? 1024 : 768) / THREADS)
env_step_kernel(const DdqcEnvParams P, const DdqcEnvState S, const DdqcEnvOut O, DdqcEnvCtl* __restrict__ ctl,
                const float* __restrict__ actions, const float* __restrict__ meas_override, const int n) {
  constexpr int A = (ACT == DDQC_ACT_CTBR_FIXED_YAW) ? 3 : 4;
  constexpr int NOBS = 13 + A;
  constexpr bool USES_CTBR = (ACT != DDQC_ACT_ROTOR_RPMS);
  const bool need_euler = !LEAN && (P.flags & DDQC_F_NEED_EULER);
  const float act_noise_std = LEAN ? 0.0f : P.actuator_noise_std;
  const float obs_noise_std = LEAN ? 0.0f : P.obs_noise_std;

  // Threads past the end (last CTA only) recompute env n-1 and skip every side effect, so the
  // body is straight-line code and the warp-level reductions below need no special casing.
  const int gid = blockIdx.x * THREADS + threadIdx.x;
  const bool live = gid < n;
  const int i = live ? gid : n - 1;
  const int lane = threadIdx.x & 31;
  const int copy = (blockIdx.x + (threadIdx.x >> 5)) % DDQC_STAT_COPIES;
  const bool indep = (P.flags & DDQC_F_INDEPENDENT_ENVS) != 0u;  // per-env instead of batch-global reset side effects
  if (!FUSED) pdl_launch_dependents();  // the finalisation may be scheduled; it waits for this grid to complete

  float rew_total = 0.0f;
  bool did_reset = false, did_timeout = false, below_ground = false;

  {
    // ------------------------------------------------------------------ load state
    V3 p{ld_state(&S.pos[i]), ld_state(&S.pos[n + i]), ld_state(&S.pos[2 * n + i])};
    Q4 q{ld_state(&S.quat[i]), ld_state(&S.quat[n + i]), ld_state(&S.quat[2 * n + i]), ld_state(&S.quat[3 * n + i])};
    V3 v{ld_state(&S.vel[i]), ld_state(&S.vel[n + i]), ld_state(&S.vel[2 * n + i])};
    V3 w{ld_state(&S.ang[i]), ld_state(&S.ang[n + i]), ld_state(&S.ang[2 * n + i])};
    V3 cmd{ld_state(&S.commands[i]), ld_state(&S.commands[n + i]), ld_state(&S.commands[2 * n + i])};
    float last_act[A], act[A];
#pragma unroll
    for (int j = 0; j < A; ++j) {
      last_act[j] = ld_state(&S.last_actions[j * n + i]);
      act[j] = clipf(actions[(size_t)i * A + j], -P.clip_actions, P.clip_actions);  // A1
    }
    int ep_len = ld_state(&S.episode_length[i]);
    int hover = ld_state(&S.hover_counter[i]);
    float sums[DDQC_NUM_REWARDS];
#pragma unroll
    for (int k = 2; k < DDQC_NUM_REWARDS; ++k) sums[k] = ld_state(&S.episode_sums[k * n + i]);

    // Everything above was written by the previous STEP kernel (or earlier), which had completed before the previous
    // finalisation released this grid (it triggers its dependents after its own wait), so those loads are in flight
    // while the finalisation still runs.  What the finalisation itself writes is read below the wait: the control
    // block and the two episode sums its A15/A16 fix-ups may touch.
    if (!FUSED) pdl_wait();
    const uint64_t ev = ctl->rng_event;
    const bool pid_pending = !indep && ctl->pid_reset_pending != 0u;
    sums[0] = ld_state(&S.episode_sums[i]);
    sums[1] = ld_state(&S.episode_sums[n + i]);

    // ------------------------------------------------------------------ action -> rpm
    float rpm[4];
    {
      float y[A];
#pragma unroll
      for (int j = 0; j < A; ++j) {
        float ex = (P.flags & DDQC_F_ACTION_LATENCY) ? last_act[j] : act[j];  // A2
        y[j] = P.act_lo[j] + ((ex - (-1.0f)) * P.act_span[j]) / 2.0f;         // A3
      }
      if (USES_CTBR) {
        float integ[3] = {0.0f, 0.0f, 0.0f}, prev[3] = {0.0f, 0.0f, 0.0f};
        if (!pid_pending) {  // A17: a reset anywhere in the previous launch zeroed every PID
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            integ[j] = ld_state(&S.pid_integral[j * n + i]);
            prev[j] = ld_state(&S.pid_prev_error[j * n + i]);
          }
        }
        V3 meas;
        if (!LEAN && meas_override) {
          meas = V3{meas_override[i], meas_override[n + i], meas_override[2 * n + i]};
        } else {  // A5: body rates as refreshed at the end of the previous step
          Q4 qc{q.w, -q.x, -q.y, -q.z};
          meas = rotate_by_quat(rotate_by_quat(w, q), qc);
        }
        float sp[3];
        sp[0] = y[1];
        sp[1] = y[2];
        sp[2] = (ACT == DDQC_ACT_CTBR) ? y[A - 1] : 0.0f;  // A4
        float ms[3] = {meas.x, meas.y, meas.z};
        float c[4];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float err = sp[j] - ms[j];
          integ[j] = integ[j] + err * P.step_dt;  // A6
          float deriv = (err - prev[j]) * P.inv_step_dt;
          prev[j] = err;
          c[j] = (P.pid_kp[j] * err + P.pid_ki[j] * integ[j]) + P.pid_kd[j] * deriv;
          if (live) {
            st_state(&S.pid_integral[j * n + i], integ[j]);
            st_state(&S.pid_prev_error[j * n + i], prev[j]);
          }
        }
        c[3] = y[0];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float f = ((P.mixer[4 * r + 0] * c[0] + P.mixer[4 * r + 1] * c[1]) + P.mixer[4 * r + 2] * c[2]) +
                    P.mixer[4 * r + 3] * c[3];
          f = f < 0.0f ? 0.0f : f;  // A7
          rpm[r] = sqrt_nonneg(f) * P.inv_sqrt_kf;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) rpm[r] = y[r < A ? r : 0];
      }
    }
    if (act_noise_std > 0.0f) {
      float z[4];
      philox_normals<4>((uint32_t)i, ev, DDQC_STREAM_ACT_NOISE, P.seed, z);
#pragma unroll
      for (int r = 0; r < 4; ++r) rpm[r] = fminf(fmaxf(rpm[r] + z[r] * act_noise_std, P.min_rpm), P.max_rpm);
    }
    if (!LEAN && live && O.rpm) {
#pragma unroll
      for (int r = 0; r < 4; ++r) O.rpm[r * n + i] = rpm[r];
    }

    // ------------------------------------------------------------------ physics (A8)
    const V3 last_pos = p;
    {
      const Wrench W = rotor_wrench(P, rpm);  // zero-order hold: constant over the substeps
      for (int sidx = 0; sidx < P.decimation; ++sidx) substep(P, W, p, q, v, w);
    }

    // ------------------------------------------------------------------ buffers (A9-A12)
    ep_len += 1;
    V3 rel{cmd.x - p.x, cmd.y - p.y, cmd.z - p.z};
    V3 last_rel{cmd.x - last_pos.x, cmd.y - last_pos.y, cmd.z - last_pos.z};
    V3 euler{0.0f, 0.0f, 0.0f};
    if (need_euler) {
      Q4 r = quat_mul(q, Q4{P.inv_init_quat[0], P.inv_init_quat[1], P.inv_init_quat[2], P.inv_init_quat[3]});
      const float rad2deg = 57.29577951308232f;
      euler.x = atan2f((r.w * r.x + r.y * r.z) * 2.0f, 1.0f - (r.x * r.x + r.y * r.y) * 2.0f) * rad2deg;
      euler.y = asinf(clipf((r.w * r.y - r.z * r.x) * 2.0f, -1.0f, 1.0f)) * rad2deg;
      euler.z = atan2f((r.w * r.z + r.x * r.y) * 2.0f, 1.0f - (r.y * r.y + r.z * r.z) * 2.0f) * rad2deg;
    }
    Q4 qc{q.w, -q.x, -q.y, -q.z};
    V3 lin_b = rotate_by_quat(v, qc);
    V3 ang_b = rotate_by_quat(rotate_by_quat(w, q), qc);

    // ------------------------------------------------------------------ hover / resample (A13-A15)
    bool hover_resampled = false;
    V3 cmd_new = cmd;
    if (P.flags & DDQC_F_AUTO_TARGET) {
      bool at_target = sqrt_nonneg(sumsq3(rel)) < P.at_target_threshold;
      bool fire;
      if (P.min_hover_steps > 0) {
        hover = at_target ? hover + 1 : 0;
        fire = hover >= P.min_hover_steps;
      } else {
        fire = at_target;
      }
      if (fire) {
        cmd_new = resample_command(P, (uint32_t)i, ev, DDQC_STREAM_HOVER);
        hover = 0;
        hover_resampled = true;
      }
    }

    // ------------------------------------------------------------------ termination (A8, A9)
    bool crash = (fabsf(euler.y) > P.term_pitch) | (fabsf(euler.x) > P.term_roll) | (fabsf(rel.x) > P.term_x) |
                 (fabsf(rel.y) > P.term_y) | (fabsf(rel.z) > P.term_z) | (p.z < P.term_ground) |
                 (fabsf(ang_b.x) > P.term_ang_vel) | (fabsf(ang_b.y) > P.term_ang_vel) |
                 (fabsf(ang_b.z) > P.term_ang_vel) | (fabsf(lin_b.x) > P.term_lin_vel) |
                 (fabsf(lin_b.y) > P.term_lin_vel) | (fabsf(lin_b.z) > P.term_lin_vel);
    below_ground = live & (p.z < 0.0f);
    did_timeout = live & (ep_len > P.max_episode_length);
    did_reset = live & (did_timeout | crash);

    float last_act_obs[A];
#pragma unroll
    for (int j = 0; j < A; ++j) last_act_obs[j] = last_act[j];
    Q4 q_obs = q;

  // A18: sum of episode_sums[k] over the envs that reset, reduced per warp without divergence --
  // for every reset lane L its 7 sums are broadcast and lane k keeps the running total of term k;
  // lanes 0..6 then issue ONE 7-address RED per warp (no shared memory, no CTA barrier).
  const unsigned int rmask = __ballot_sync(0xffffffffu, did_reset);
  if (rmask) {
    float part = 0.0f;
    for (unsigned int m = rmask; m; m &= m - 1) {
      const int L = __ffs(m) - 1;
#pragma unroll
      for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
        float vk = __shfl_sync(0xffffffffu, sums[k], L);
        if (lane == k) part += vk;
      }
    }
    if (lane < DDQC_NUM_REWARDS) atomicAdd(&ctl->acc_sums[copy][lane], part);
    if (lane == 8) atomicAdd(&ctl->acc_reset[copy], (unsigned int)__popc(rmask));
  }
  if (did_reset) {
#pragma unroll
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) sums[k] = 0.0f;
  }

    if (did_reset) {  // reset_idx (hover_env.py:556-595)
      if (USES_CTBR && indep && live) {  // ctbr_controller.reset() of this env's own instance
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          S.pid_integral[j * n + i] = 0.0f;
          S.pid_prev_error[j * n + i] = 0.0f;
        }
      }
      V3 ip{P.init_pos[0], P.init_pos[1], P.init_pos[2]};
      p = ip;
      rel = V3{cmd_new.x - ip.x, cmd_new.y - ip.y, cmd_new.z - ip.z};  // A16, old (pre-reset-resample) command
      last_rel = rel;
      q = Q4{P.init_quat[0], P.init_quat[1], P.init_quat[2], P.init_quat[3]};
      q_obs = q;
      v = V3{0.0f, 0.0f, 0.0f};
      w = V3{0.0f, 0.0f, 0.0f};
      lin_b = V3{0.0f, 0.0f, 0.0f};
      ang_b = V3{0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int j = 0; j < A; ++j) last_act_obs[j] = 0.0f;
      ep_len = 0;
      hover = 0;
      cmd_new = resample_command(P, (uint32_t)i, ev, DDQC_STREAM_RESET);
      hover_resampled = false;
    }

    // ------------------------------------------------------------------ rewards (A19)
    float r_target = sumsq3(last_rel) - sumsq3(rel);
    float r_close = (P.at_target_threshold_sq - sumsq3(rel)) * P.inv_at_target_threshold_sq;
    r_close = r_close < 0.0f ? 0.0f : r_close;
    float r_hover = (float)hover;
    float r_smooth;
    {
      float d0 = act[0] - last_act_obs[0];
      r_smooth = d0 * d0;
#pragma unroll
      for (int j = 1; j < A; ++j) {
        float d = act[j] - last_act_obs[j];
        r_smooth = r_smooth + d * d;
      }
    }
    float r_yaw = 0.0f;
    if (need_euler) {
      float yaw = euler.z;
      yaw = (yaw > 180.0f ? yaw - 360.0f : yaw) * P.inv_180 * 3.14159f;
      r_yaw = expf(P.yaw_lambda * fabsf(yaw));
    }
    float r_ang = sqrt_nonneg(sumsq3(V3{ang_b.x * P.inv_pi_lit, ang_b.y * P.inv_pi_lit, ang_b.z * P.inv_pi_lit}));
    float r_crash = crash ? 1.0f : 0.0f;

    float rk[DDQC_NUM_REWARDS] = {r_target * P.rew_scale[0], r_close * P.rew_scale[1], r_hover * P.rew_scale[2],
                                  r_smooth * P.rew_scale[3], r_yaw * P.rew_scale[4],  r_ang * P.rew_scale[5],
                                  r_crash * P.rew_scale[6]};
    const float sum0_pre = sums[0], sum1_pre = sums[1];
    float rew = 0.0f;
#pragma unroll
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
      rew = rew + rk[k];
      sums[k] = sums[k] + rk[k];
    }
    rew_total = live ? rew : 0.0f;

    // ------------------------------------------------------------------ observations (A20)
    float obs[NOBS];
    obs[0] = clipf(rel.x * P.obs_scale_pos, -1.0f, 1.0f);
    obs[1] = clipf(rel.y * P.obs_scale_pos, -1.0f, 1.0f);
    obs[2] = clipf(rel.z * P.obs_scale_pos, -1.0f, 1.0f);
    obs[3] = q_obs.w;
    obs[4] = q_obs.x;
    obs[5] = q_obs.y;
    obs[6] = q_obs.z;
    obs[7] = clipf(lin_b.x * P.obs_scale_lin, -1.0f, 1.0f);
    obs[8] = clipf(lin_b.y * P.obs_scale_lin, -1.0f, 1.0f);
    obs[9] = clipf(lin_b.z * P.obs_scale_lin, -1.0f, 1.0f);
    obs[10] = clipf(ang_b.x * P.obs_scale_ang, -1.0f, 1.0f);
    obs[11] = clipf(ang_b.y * P.obs_scale_ang, -1.0f, 1.0f);
    obs[12] = clipf(ang_b.z * P.obs_scale_ang, -1.0f, 1.0f);
#pragma unroll
    for (int j = 0; j < A; ++j) obs[13 + j] = last_act_obs[j];
    if (obs_noise_std > 0.0f) {
      float z[13];
      philox_normals<13>((uint32_t)i, ev, DDQC_STREAM_OBS_NOISE, P.seed, z);
      const float s = obs_noise_std;
      obs[0] = obs[0] + z[0] * s;
      obs[1] = obs[1] + z[1] * s;
      obs[2] = obs[2] + z[2] * s;
      obs[7] = obs[7] + z[3] * s;
      obs[8] = obs[8] + z[4] * s;
      obs[9] = obs[9] + z[5] * s;
      obs[10] = obs[10] + z[6] * s;
      obs[11] = obs[11] + z[7] * s;
      obs[12] = obs[12] + z[8] * s;
      float a = obs[3] + z[9] * s, b = obs[4] + z[10] * s, c = obs[5] + z[11] * s, d = obs[6] + z[12] * s;
      float nrm = sqrtf(((a * a + b * b) + c * c) + d * d);
      obs[3] = a / nrm;
      obs[4] = b / nrm;
      obs[5] = c / nrm;
      obs[6] = d / nrm;
    }

    // A15/A16 fix-up record: the values this env would have if some other env reset this step
    if (hover_resampled && live) {
      V3 rel_f{cmd_new.x - p.x, cmd_new.y - p.y, cmd_new.z - p.z};
      V3 lrel_f{cmd_new.x - last_pos.x, cmd_new.y - last_pos.y, cmd_new.z - last_pos.z};
      float rt = (sumsq3(lrel_f) - sumsq3(rel_f)) * P.rew_scale[0];
      float rc = (P.at_target_threshold_sq - sumsq3(rel_f)) * P.inv_at_target_threshold_sq;
      rc = (rc < 0.0f ? 0.0f : rc) * P.rew_scale[1];
      float rw = 0.0f + rt;
      rw = rw + rc;
#pragma unroll
      for (int k = 2; k < DDQC_NUM_REWARDS; ++k) rw = rw + rk[k];
      unsigned int slot = atomicAdd(&ctl->fixup_count, 1u);
      O.fixup_idx[slot] = i;
      float* fv = O.fixup_val + (size_t)slot * 8;
      fv[0] = rw;
      fv[1] = sum0_pre + rt;
      fv[2] = sum1_pre + rc;
      fv[3] = clipf(rel_f.x * P.obs_scale_pos, -1.0f, 1.0f);
      fv[4] = clipf(rel_f.y * P.obs_scale_pos, -1.0f, 1.0f);
      fv[5] = clipf(rel_f.z * P.obs_scale_pos, -1.0f, 1.0f);
      fv[6] = 0.0f;
      fv[7] = 0.0f;
    }

    // ------------------------------------------------------------------ store
    if (live) {
    st_state(&S.pos[i], p.x);
    st_state(&S.pos[n + i], p.y);
    st_state(&S.pos[2 * n + i], p.z);
    st_state(&S.quat[i], q.w);
    st_state(&S.quat[n + i], q.x);
    st_state(&S.quat[2 * n + i], q.y);
    st_state(&S.quat[3 * n + i], q.z);
    st_state(&S.vel[i], v.x);
    st_state(&S.vel[n + i], v.y);
    st_state(&S.vel[2 * n + i], v.z);
    st_state(&S.ang[i], w.x);
    st_state(&S.ang[n + i], w.y);
    st_state(&S.ang[2 * n + i], w.z);
    st_state(&S.commands[i], cmd_new.x);
    st_state(&S.commands[n + i], cmd_new.y);
    st_state(&S.commands[2 * n + i], cmd_new.z);
#pragma unroll
    for (int j = 0; j < A; ++j) st_state(&S.last_actions[j * n + i], act[j]);  // hover_env.py:544, all envs
    st_state(&S.episode_length[i], ep_len);
    st_state(&S.hover_counter[i], hover);
#pragma unroll
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) st_state(&S.episode_sums[k * n + i], sums[k]);

    if (NOBS == 16) {
      float4* o4 = reinterpret_cast<float4*>(O.obs + (size_t)i * 16);
      st_out(&o4[0], make_float4(obs[0], obs[1], obs[2], obs[3]));
      st_out(&o4[1], make_float4(obs[4], obs[5], obs[6], obs[7]));
      st_out(&o4[2], make_float4(obs[8], obs[9], obs[10], obs[11]));
      st_out(&o4[3], make_float4(obs[12], obs[13], obs[14], obs[15]));
    } else {
#pragma unroll
      for (int j = 0; j < NOBS; ++j) O.obs[(size_t)i * NOBS + j] = obs[j];
    }
    st_out(&O.rew[i], rew);
    st_out(&O.reset[i], (uint8_t)(did_reset ? 1 : 0));
    st_out(&O.time_outs[i], did_timeout ? 1.0f : 0.0f);

    if (!LEAN && O.base_lin_vel) {
      O.base_lin_vel[i] = lin_b.x;
      O.base_lin_vel[n + i] = lin_b.y;
      O.base_lin_vel[2 * n + i] = lin_b.z;
    }
    if (!LEAN && O.base_ang_vel) {
      O.base_ang_vel[i] = ang_b.x;
      O.base_ang_vel[n + i] = ang_b.y;
      O.base_ang_vel[2 * n + i] = ang_b.z;
    }
    if (!LEAN && O.rel_pos) {
      O.rel_pos[i] = rel.x;
      O.rel_pos[n + i] = rel.y;
      O.rel_pos[2 * n + i] = rel.z;
    }
    if (!LEAN && O.last_rel_pos) {
      O.last_rel_pos[i] = last_rel.x;
      O.last_rel_pos[n + i] = last_rel.y;
      O.last_rel_pos[2 * n + i] = last_rel.z;
    }
    if (!LEAN && O.base_euler) {
      O.base_euler[i] = euler.x;
      O.base_euler[n + i] = euler.y;
      O.base_euler[2 * n + i] = euler.z;
    }
    if (!LEAN && O.crash) O.crash[i] = crash ? 1 : 0;
    if (!LEAN && O.last_base_pos) {
      O.last_base_pos[i] = did_reset ? P.init_pos[0] : last_pos.x;
      O.last_base_pos[n + i] = did_reset ? P.init_pos[1] : last_pos.y;
      O.last_base_pos[2 * n + i] = did_reset ? P.init_pos[2] : last_pos.z;
    }
    }  // live
  }

  // -------------------------------------------------------------------- statistics
  // time-out count by ballot, reward total by a shuffle tree; one RED per warp each
  {
    unsigned int tmask = __ballot_sync(0xffffffffu, did_timeout);
    float r = rew_total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 7) atomicAdd(&ctl->acc_sums[copy][7], r);
    if (lane == 9 && tmask) atomicAdd(&ctl->acc_timeout[copy], (unsigned int)__popc(tmask));
    if (P.flags & DDQC_F_GROUND_WATCH) {  // no ground plane here: count the env-steps the reference's floor would have changed
      const unsigned int gmask = __ballot_sync(0xffffffffu, below_ground);
      if (lane == 10 && gmask) atomicAdd(&ctl->acc_ground[copy], (unsigned int)__popc(gmask));
    }
  }

  if (FUSED) {
    // ------------------------------------------------------------------ last-CTA finalisation
    // bar.sync orders every warp's stores / REDs before thread 0's device-scope fence (cumulative),
    // which releases them to whichever CTA draws the last ticket.  Only warp 0 stays for the ticket.
    __syncthreads();
    if (threadIdx.x >= 32) return;
    int is_last = 0;
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = (atomicAdd(&ctl->blocks_done, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    is_last = __shfl_sync(0xffffffffu, is_last, 0);
    if (!is_last) return;
    __threadfence();
    finalize_step<NOBS, LEAN, USES_CTBR>(P, S, O, ctl, n, lane);
  }
}

// Second launch of a step: one warp folds the statistics of the step kernel that has just completed
// (the kernel boundary is the only ordering the stores, the REDs and the fix-up list need: the step
// kernel itself has no barrier, no fence and no ticket, so a CTA's registers are released the moment
// its last store is issued).
template <int NOBS, bool LEAN, bool USES_CTBR>
__global__ void __launch_bounds__(32, 1)
env_finalize_kernel(const DdqcEnvParams P, const DdqcEnvState S, const DdqcEnvOut O, DdqcEnvCtl* __restrict__ ctl,
                    const int n) {
  pdl_wait();               // the step kernel has completed and its stores / REDs are visible
  pdl_launch_dependents();  // only now: the next step kernel loads state the completed step kernel wrote before ITS wait
  finalize_step<NOBS, LEAN, USES_CTBR>(P, S, O, ctl, n, (int)threadIdx.x);
}

}  // namespace

static int check_env_args(const DdqcEnvParams* P, const DdqcEnvState* S, int64_t n) {
  if (!P || !S) return DDQC_E_NULL;
  if (n <= 0 || n > 0x7fffffff / 8) return DDQC_E_SHAPE;
  if (P->action_type < 0 || P->action_type > 2) return DDQC_E_CONFIG;
  if (P->decimation < 1) return DDQC_E_CONFIG;
  if (!S->pos || !S->quat || !S->vel || !S->ang || !S->commands || !S->last_actions || !S->episode_sums ||
      !S->episode_length || !S->hover_counter)
    return DDQC_E_NULL;
  if (P->action_type != DDQC_ACT_ROTOR_RPMS && (!S->pid_integral || !S->pid_prev_error)) return DDQC_E_NULL;
  return 0;
}

extern "C" int ddqc_env_step(const DdqcEnvParams* P, const DdqcEnvState* S, const DdqcEnvOut* O, DdqcEnvCtl* ctl,
                             const float* actions, const float* meas_override, int64_t n_env, void* stream) {
  int rc = check_env_args(P, S, n_env);
  if (rc) return rc;
  if (!O || !ctl || !actions || !O->obs || !O->rew || !O->reset || !O->time_outs || !O->fixup_idx || !O->fixup_val)
    return DDQC_E_NULL;
  if (P->action_type == DDQC_ACT_CTBR_FIXED_YAW && (reinterpret_cast<uintptr_t>(O->obs) & 15u)) return DDQC_E_ALIGN;
  const int n = (int)n_env;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool lean = !(P->flags & DDQC_F_NEED_EULER) && !(P->actuator_noise_std > 0.0f) && !(P->obs_noise_std > 0.0f) &&
                    !meas_override && !O->base_lin_vel && !O->base_ang_vel && !O->rel_pos && !O->last_rel_pos &&
                    !O->base_euler && !O->crash && !O->rpm && !O->last_base_pos;
  static const int64_t fused_max_n = [] {
    const char* e = getenv("DDQC_ENV_FUSED_MAX_N");
    return e ? (int64_t)atoll(e) : (int64_t)DDQC_ENV_FUSED_MAX_N;
  }();
  const bool fused = n_env <= fused_max_n;
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaError_t err = cudaSuccess;
  auto launch = [&](auto kernel, dim3 g, dim3 b, auto... args) {
    if (err != cudaSuccess) return;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = g;
    cfg.blockDim = b;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = (DDQC_ENV_FINALIZE == 2 && !fused) ? 1 : 0;
    err = cudaLaunchKernelEx(&cfg, kernel, args...);
  };
#define DDQC_LAUNCH2(ACT, LEANV)                                                                              \
  do {                                                                                                        \
    if (fused) {                                                                                              \
      launch(env_step_kernel<ACT, LEANV, true, kThreadsFused>, dim3((n + kThreadsFused - 1) / kThreadsFused),  \
             dim3(kThreadsFused), *P, *S, *O, ctl, actions, meas_override, n);                                \
    } else {                                                                                                  \
      launch(env_step_kernel<ACT, LEANV, false, kThreadsSplit>, dim3((n + kThreadsSplit - 1) / kThreadsSplit), \
             dim3(kThreadsSplit), *P, *S, *O, ctl, actions, meas_override, n);                                \
      launch(env_finalize_kernel<(ACT == DDQC_ACT_CTBR_FIXED_YAW ? 16 : 17), LEANV, (ACT != DDQC_ACT_ROTOR_RPMS)>, \
             dim3(1), dim3(32), *P, *S, *O, ctl, n);                                                          \
    }                                                                                                         \
  } while (0)
#define DDQC_LAUNCH(ACT)                                                                                      \
  do {                                                                                                        \
    if (lean)                                                                                                 \
      DDQC_LAUNCH2(ACT, true);                                                                                \
    else                                                                                                      \
      DDQC_LAUNCH2(ACT, false);                                                                               \
  } while (0)
  switch (P->action_type) {
    case DDQC_ACT_ROTOR_RPMS:
      DDQC_LAUNCH(DDQC_ACT_ROTOR_RPMS);
      break;
    case DDQC_ACT_CTBR:
      DDQC_LAUNCH(DDQC_ACT_CTBR);
      break;
    default:
      DDQC_LAUNCH(DDQC_ACT_CTBR_FIXED_YAW);
      break;
  }
#undef DDQC_LAUNCH
#undef DDQC_LAUNCH2
  return (int)err;
}
