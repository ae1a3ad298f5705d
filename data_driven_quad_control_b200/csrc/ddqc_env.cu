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
#include <string.h>

#include "ddqc_device.cuh"
#include "ddqc_env_pair.cuh"

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
// ... and the bulk-copy tile kernel (env_step_tile_kernel) from DDQC_ENV_TILE_MIN_N envs on, in the lean configuration
// (environment variable of the same name overrides; a huge value switches the tile kernel off)
#ifndef DDQC_ENV_TILE_MIN_N
#define DDQC_ENV_TILE_MIN_N (1ll << 40)  // off: measured slower than the plain kernel (see env_step_tile_kernel)
#endif
// ... and the pair kernel (two envs per thread on packed fp32x2 arithmetic) from DDQC_ENV_PAIR_MIN_N envs on
#ifndef DDQC_ENV_PAIR_MIN_N
#define DDQC_ENV_PAIR_MIN_N (1ll << 40)  // off: 28 % fewer instructions, but 16 warps per SM instead of 32 - 69.7 us against 64.3 us
#endif
// ... and the pair kernel on bulk-copy staged inputs (env_step_pair_tile_kernel) from DDQC_ENV_PAIR_TILE_MIN_N envs on
#ifndef DDQC_ENV_PAIR_TILE_MIN_N
#define DDQC_ENV_PAIR_TILE_MIN_N (1ll << 40)  // off: 78.0 us
#endif
#ifndef DDQC_ENV_TILE
#define DDQC_ENV_TILE 256  // envs per tile of the tile kernel = threads per group (1024 threads per CTA)
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

// ---- one env-step for one env ------------------------------------------------------------------
// The whole of HoverEnv.step for env `i`, state in registers from the first load to the last store.  `IO` says where
// the env's struct-of-arrays state and its per-env outputs live: GlobalIO reads and writes global memory directly
// (one coalesced 32-bit access per row and thread), TileIO a shared-memory tile that bulk copies fill and drain.
enum Field { F_POS, F_QUAT, F_VEL, F_ANG, F_CMD, F_LACT, F_EPLEN, F_HOVER, F_SUMS, F_PIDI, F_PIDE };

template <int ACT, bool LEAN, class IO>
__device__ __forceinline__ void env_step_body(const DdqcEnvParams& P, const DdqcEnvOut& O, DdqcEnvCtl* __restrict__ ctl, IO& io,
                                              const float* __restrict__ meas_override, const int n, const int i,
                                              const bool live, const int lane, const int copy, float& rew_total,
                                              bool& did_timeout, bool& below_ground) {
  constexpr int A = (ACT == DDQC_ACT_CTBR_FIXED_YAW) ? 3 : 4;
  constexpr int NOBS = 13 + A;
  constexpr bool USES_CTBR = (ACT != DDQC_ACT_ROTOR_RPMS);
  const bool need_euler = !LEAN && (P.flags & DDQC_F_NEED_EULER);
  const float act_noise_std = LEAN ? 0.0f : P.actuator_noise_std;
  const float obs_noise_std = LEAN ? 0.0f : P.obs_noise_std;
  const bool indep = (P.flags & DDQC_F_INDEPENDENT_ENVS) != 0u;  // per-env instead of batch-global reset side effects
  bool did_reset = false;
  io.begin();
  {
    // ------------------------------------------------------------------ load state
    V3 p{io.template ld<F_POS>(0), io.template ld<F_POS>(1), io.template ld<F_POS>(2)};
    Q4 q{io.template ld<F_QUAT>(0), io.template ld<F_QUAT>(1), io.template ld<F_QUAT>(2), io.template ld<F_QUAT>(3)};
    V3 v{io.template ld<F_VEL>(0), io.template ld<F_VEL>(1), io.template ld<F_VEL>(2)};
    V3 w{io.template ld<F_ANG>(0), io.template ld<F_ANG>(1), io.template ld<F_ANG>(2)};
    V3 cmd{io.template ld<F_CMD>(0), io.template ld<F_CMD>(1), io.template ld<F_CMD>(2)};
    float last_act[A], act[A];
#pragma unroll
    for (int j = 0; j < A; ++j) {
      last_act[j] = io.template ld<F_LACT>(j);
      act[j] = clipf(io.action(j), -P.clip_actions, P.clip_actions);  // A1
    }
    int ep_len = io.template ldi<F_EPLEN>();
    int hover = io.template ldi<F_HOVER>();
    float sums[DDQC_NUM_REWARDS];
#pragma unroll
    for (int k = 2; k < DDQC_NUM_REWARDS; ++k) sums[k] = io.template ld<F_SUMS>(k);

    // Everything above was written by the previous STEP kernel (or earlier), which had completed before the previous
    // finalisation released this grid (it triggers its dependents after its own wait), so those loads are in flight
    // while the finalisation still runs.  What the finalisation itself writes is read below the wait: the control
    // block and the two episode sums its A15/A16 fix-ups may touch.
    io.late_sync();
    const uint64_t ev = io.rng_event(ctl);
    const bool pid_pending = !indep && io.pid_reset_pending(ctl);
    sums[0] = io.template ld<F_SUMS>(0);
    sums[1] = io.template ld<F_SUMS>(1);
    float integ[3] = {0.0f, 0.0f, 0.0f}, prev[3] = {0.0f, 0.0f, 0.0f};
    // A17: a reset anywhere in the previous launch zeroed every PID.  A shared-memory tile must be read completely before
    // it is refilled (IO::kReadAllFirst); from global memory the loads sit where the values are used - hoisting them
    // here costs the 64-register kernel 0.9 us per 2^20 envs
    if (IO::kReadAllFirst) {
      if (USES_CTBR && !pid_pending) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          integ[j] = io.template ld<F_PIDI>(j);
          prev[j] = io.template ld<F_PIDE>(j);
        }
      }
      io.inputs_done();  // nothing of this env's stored state is read below this line
    }

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
        if (!IO::kReadAllFirst && !pid_pending) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            integ[j] = io.template ld<F_PIDI>(j);
            prev[j] = io.template ld<F_PIDE>(j);
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
            io.template st<F_PIDI>(j, integ[j]);
            io.template st<F_PIDE>(j, prev[j]);
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
      substeps(P, W, p, q, v, w);
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
          io.template st<F_PIDI>(j, 0.0f);
          io.template st<F_PIDE>(j, 0.0f);
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
    io.template st<F_POS>(0, p.x);
    io.template st<F_POS>(1, p.y);
    io.template st<F_POS>(2, p.z);
    io.template st<F_QUAT>(0, q.w);
    io.template st<F_QUAT>(1, q.x);
    io.template st<F_QUAT>(2, q.y);
    io.template st<F_QUAT>(3, q.z);
    io.template st<F_VEL>(0, v.x);
    io.template st<F_VEL>(1, v.y);
    io.template st<F_VEL>(2, v.z);
    io.template st<F_ANG>(0, w.x);
    io.template st<F_ANG>(1, w.y);
    io.template st<F_ANG>(2, w.z);
    io.template st<F_CMD>(0, cmd_new.x);
    io.template st<F_CMD>(1, cmd_new.y);
    io.template st<F_CMD>(2, cmd_new.z);
#pragma unroll
    for (int j = 0; j < A; ++j) io.template st<F_LACT>(j, act[j]);  // hover_env.py:544, all envs
    io.template sti<F_EPLEN>(ep_len);
    io.template sti<F_HOVER>(hover);
#pragma unroll
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) io.template st<F_SUMS>(k, sums[k]);

    if (NOBS == 16) {  // 64-byte observation rows: four 16-byte stores straight to global memory in both variants
      float4* o4 = reinterpret_cast<float4*>(O.obs + (size_t)i * 16);
      o4[0] = make_float4(obs[0], obs[1], obs[2], obs[3]);
      o4[1] = make_float4(obs[4], obs[5], obs[6], obs[7]);
      o4[2] = make_float4(obs[8], obs[9], obs[10], obs[11]);
      o4[3] = make_float4(obs[12], obs[13], obs[14], obs[15]);
    } else {
#pragma unroll
      for (int j = 0; j < NOBS; ++j) O.obs[(size_t)i * NOBS + j] = obs[j];
    }
    io.rew(rew);
    io.reset((uint8_t)(did_reset ? 1 : 0));
    io.time_out(did_timeout ? 1.0f : 0.0f);

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

}

// time-out count by ballot, reward total by a shuffle tree, below-ground count by ballot; one RED per warp each
__device__ __forceinline__ void warp_stats(const DdqcEnvParams& P, DdqcEnvCtl* __restrict__ ctl, const int lane, const int copy,
                                           const float rew_total, const bool did_timeout, const bool below_ground) {
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

// state and outputs in global memory
template <int A, bool FUSED>
struct GlobalIO {
  static constexpr bool kReadAllFirst = false;
  const DdqcEnvState& S;
  const DdqcEnvOut& O;
  const float* __restrict__ actions;
  const int n, i;
  const bool live;
  template <int F>
  __device__ __forceinline__ float* fptr() const {
    if constexpr (F == F_POS) return S.pos;
    else if constexpr (F == F_QUAT) return S.quat;
    else if constexpr (F == F_VEL) return S.vel;
    else if constexpr (F == F_ANG) return S.ang;
    else if constexpr (F == F_CMD) return S.commands;
    else if constexpr (F == F_LACT) return S.last_actions;
    else if constexpr (F == F_SUMS) return S.episode_sums;
    else if constexpr (F == F_PIDI) return S.pid_integral;
    else return S.pid_prev_error;
  }
  template <int F>
  __device__ __forceinline__ float ld(int c) const { return ld_state(&fptr<F>()[c * n + i]); }
  template <int F>
  __device__ __forceinline__ int ldi() const { return ld_state(F == F_EPLEN ? &S.episode_length[i] : &S.hover_counter[i]); }
  template <int F>
  __device__ __forceinline__ void st(int c, float v) const { fptr<F>()[c * n + i] = v; }  // callers guard with `live`
  template <int F>
  __device__ __forceinline__ void sti(int v) const { (F == F_EPLEN ? S.episode_length : S.hover_counter)[i] = v; }
  __device__ __forceinline__ float action(int j) const { return actions[(size_t)i * A + j]; }
  __device__ __forceinline__ void rew(float v) const { O.rew[i] = v; }
  __device__ __forceinline__ void reset(uint8_t v) const { O.reset[i] = v; }
  __device__ __forceinline__ void time_out(float v) const { O.time_outs[i] = v; }
  __device__ __forceinline__ void begin() const {}
  __device__ __forceinline__ void inputs_done() const {}
  __device__ __forceinline__ uint64_t rng_event(const DdqcEnvCtl* ctl) const { return ctl->rng_event; }
  __device__ __forceinline__ bool pid_reset_pending(const DdqcEnvCtl* ctl) const { return ctl->pid_reset_pending != 0u; }
  // What the finalisation of the previous step writes (control block, fix-up episode sums) is read after this point
  __device__ __forceinline__ void late_sync() const { if (!FUSED) pdl_wait(); }
};

// ---- the tile kernel: persistent CTAs, the state of the NEXT tile staged through shared memory by bulk copies ------
// One CTA per SM, kGroups groups of kTile threads (32 warps of 64 registers, like the plain kernel); a group walks its
// own stream of kTile-env tiles.  A tile's ~34 input rows (kTile x 4 bytes each, contiguous in the struct-of-arrays
// storage) and its action slab arrive in the group's shared-memory buffer as 1-D bulk copies (cp.async.bulk -> mbarrier
// complete_tx).  The threads read their env's column into registers, a group barrier says "the buffer has been read",
// and the lane-0 threads of the group's warps (warp w moves the rows r = w mod 8) immediately refill the SAME buffer
// with the next tile: its loads fly during the ~1100 instructions of arithmetic on the current one, so no warp ever
// waits on DRAM, and the input side needs no per-access 64-bit address arithmetic (shared-memory loads with immediate
// offsets).  The new state and the outputs go straight from the registers to global memory (stores do not stall).
// Measured and dropped on the way: outputs staged through shared memory as well and drained by bulk stores - it needs two
// 40 KB buffers per group, i.e. 16 warps per SM instead of 32, and this arithmetic (dependent fp32 chains, no FMA
// contraction) needs 8 warps per scheduler: 84.8 us per 2^20 envs against 63.4 us for the plain kernel.
constexpr int kTile = DDQC_ENV_TILE;
constexpr int kGroups = 1024 / kTile;
constexpr int kTileThreads = kTile * kGroups;
constexpr int kGroupWarps = kTile / 32;

template <int A>
struct TileRows {
  static constexpr int POS = 0, QUAT = 3, VEL = 7, ANG = 10, CMD = 13, LACT = 16, EPLEN = 16 + A, HOVER = 17 + A,
                       SUMS = 18 + A, PIDI = 25 + A, PIDE = 28 + A, NROWS = 31 + A;
  static constexpr int ROW_BYTES = kTile * 4;
  static constexpr int ACT_OFF = NROWS * kTile;                        // float index of the action slab [kTile][A]
  static constexpr int BUF_BYTES = ((NROWS + A) * ROW_BYTES + 127) / 128 * 128;
  static constexpr int BAR_OFF = BUF_BYTES * kGroups;                  // mbarriers: [group][main | late]
  static constexpr int PTR_OFF = BAR_OFF + 8 * 2 * kGroups;            // uint64 global base address of every row
  static constexpr int SMEM_BYTES = PTR_OFF + 8 * NROWS;
  template <int F>
  __host__ __device__ static constexpr int row0() {
    return F == F_POS ? POS : F == F_QUAT ? QUAT : F == F_VEL ? VEL : F == F_ANG ? ANG : F == F_CMD ? CMD : F == F_LACT ? LACT
         : F == F_EPLEN ? EPLEN : F == F_HOVER ? HOVER : F == F_SUMS ? SUMS : F == F_PIDI ? PIDI : PIDE;
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// The bulk copies of one tile into a group's buffer, all by ONE thread (the groups' warps take turns from tile to tile):
// ~5 instructions per copy, straight-line, instead of every warp of the group walking its share of the row list.  The
// expected byte counts are posted before the copies that will complete them.
template <int A, bool USES_CTBR>
__device__ __forceinline__ void tile_issue_loads(const unsigned long long* __restrict__ rowptr, const float* __restrict__ actions,
                                                 uint32_t dst0, uint32_t bar_main, int tile, bool main_rows, bool late_rows,
                                                 bool pid_rows) {
  using R = TileRows<A>;
  const uint32_t bar_late = bar_main + 8;
  const bool pid = USES_CTBR && pid_rows;
  if (main_rows) mbar_arrive_expect_tx(bar_main, (uint32_t)((R::PIDI - 2) * R::ROW_BYTES + kTile * A * 4));
  if (late_rows) mbar_arrive_expect_tx(bar_late, (uint32_t)((2 + (pid ? 6 : 0)) * R::ROW_BYTES));
  const size_t off = (size_t)tile * R::ROW_BYTES;
#pragma unroll
  for (int r = 0; r < R::NROWS; ++r) {
    const bool is_pid = r >= R::PIDI, is_late = is_pid || r == R::SUMS || r == R::SUMS + 1;
    if (is_late ? !late_rows : !main_rows) continue;
    if (is_pid && !pid) continue;
    bulk_load(dst0 + (uint32_t)(r * R::ROW_BYTES), reinterpret_cast<const char*>((uintptr_t)rowptr[r]) + off, R::ROW_BYTES,
              is_late ? bar_late : bar_main);
  }
  if (main_rows)  // the tile's action slab [kTile][A], contiguous
    bulk_load(dst0 + (uint32_t)(R::ACT_OFF * 4), actions + (size_t)tile * kTile * A, kTile * A * 4, bar_main);
}

// global base address (env 0) of input row r of the tile layout
template <int A, bool USES_CTBR>
__device__ __forceinline__ const float* tile_row_base(const DdqcEnvState& S, int r, int n) {
  using R = TileRows<A>;
  if (r < R::QUAT) return S.pos + (size_t)(r - R::POS) * n;
  if (r < R::VEL) return S.quat + (size_t)(r - R::QUAT) * n;
  if (r < R::ANG) return S.vel + (size_t)(r - R::VEL) * n;
  if (r < R::CMD) return S.ang + (size_t)(r - R::ANG) * n;
  if (r < R::LACT) return S.commands + (size_t)(r - R::CMD) * n;
  if (r < R::EPLEN) return S.last_actions + (size_t)(r - R::LACT) * n;
  if (r == R::EPLEN) return reinterpret_cast<const float*>(S.episode_length);
  if (r == R::HOVER) return reinterpret_cast<const float*>(S.hover_counter);
  if (r < R::PIDI) return S.episode_sums + (size_t)(r - R::SUMS) * n;
  if (r < R::PIDE) return USES_CTBR ? S.pid_integral + (size_t)(r - R::PIDI) * n : nullptr;
  return USES_CTBR ? S.pid_prev_error + (size_t)(r - R::PIDE) * n : nullptr;
}

// inputs of one env from the group's shared-memory buffer (column `gt` of every row), outputs to global memory
template <int A, bool USES_CTBR>
struct TileIO {
  static constexpr bool kReadAllFirst = true;
  using R = TileRows<A>;
  const float* buf;      // the group's buffer, as floats
  const DdqcEnvState& S;
  const DdqcEnvOut& O;
  const unsigned long long* rowptr;
  const float* actions;
  const int n, i, gt, g;
  const uint32_t buf_addr, bar_main, parity;
  const int next_tile;   // < 0: this stream has no further tile
  const bool issuer;     // this thread issues the next tile's copies
  const bool pid_rows;
  const uint64_t ev;     // control-block values, read once per launch
  const bool pid_pending;
  template <int F>
  __device__ __forceinline__ float* fptr() const {
    if constexpr (F == F_POS) return S.pos;
    else if constexpr (F == F_QUAT) return S.quat;
    else if constexpr (F == F_VEL) return S.vel;
    else if constexpr (F == F_ANG) return S.ang;
    else if constexpr (F == F_CMD) return S.commands;
    else if constexpr (F == F_LACT) return S.last_actions;
    else if constexpr (F == F_SUMS) return S.episode_sums;
    else if constexpr (F == F_PIDI) return S.pid_integral;
    else return S.pid_prev_error;
  }
  template <int F>
  __device__ __forceinline__ float ld(int c) const { return buf[(R::template row0<F>() + c) * kTile + gt]; }
  template <int F>
  __device__ __forceinline__ int ldi() const { return __float_as_int(buf[R::template row0<F>() * kTile + gt]); }
  template <int F>
  __device__ __forceinline__ void st(int c, float v) const { fptr<F>()[c * n + i] = v; }
  template <int F>
  __device__ __forceinline__ void sti(int v) const { (F == F_EPLEN ? S.episode_length : S.hover_counter)[i] = v; }
  __device__ __forceinline__ float action(int j) const { return buf[R::ACT_OFF + gt * A + j]; }
  __device__ __forceinline__ void rew(float v) const { O.rew[i] = v; }
  __device__ __forceinline__ void reset(uint8_t v) const { O.reset[i] = v; }
  __device__ __forceinline__ void time_out(float v) const { O.time_outs[i] = v; }
  __device__ __forceinline__ void begin() const { mbar_wait_parity(bar_main, parity); }          // the main rows have landed
  __device__ __forceinline__ void late_sync() const { mbar_wait_parity(bar_main + 8, parity); }  // ... and the late ones
  // every thread of the group has its inputs in registers: the buffer is refilled with the stream's next tile, whose
  // copies then run under the arithmetic of this one
  __device__ __forceinline__ void inputs_done() const {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kTile) : "memory");
    if (issuer && next_tile >= 0) tile_issue_loads<A, USES_CTBR>(rowptr, actions, buf_addr, bar_main, next_tile, true, true, pid_rows);
  }
  __device__ __forceinline__ uint64_t rng_event(const DdqcEnvCtl*) const { return ev; }
  __device__ __forceinline__ bool pid_reset_pending(const DdqcEnvCtl*) const { return pid_pending; }
};

template <int ACT>
__global__ void __launch_bounds__(kTileThreads, 1)
env_step_tile_kernel(const DdqcEnvParams P, const DdqcEnvState S, const DdqcEnvOut O, DdqcEnvCtl* __restrict__ ctl,
                     const float* __restrict__ actions, const int n, const int n_tiles) {
  constexpr int A = (ACT == DDQC_ACT_CTBR_FIXED_YAW) ? 3 : 4;
  constexpr bool USES_CTBR = (ACT != DDQC_ACT_ROTOR_RPMS);
  using R = TileRows<A>;
  extern __shared__ __align__(128) uint8_t tile_smem[];
  const int tid = threadIdx.x, g = tid / kTile, gt = tid % kTile, gw = gt >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(tile_smem);
  unsigned long long* rowptr = reinterpret_cast<unsigned long long*>(tile_smem + R::PTR_OFF);

  // ---- one-time setup: global base address of every input row, mbarriers ----
  if (tid < R::NROWS) rowptr[tid] = (unsigned long long)(uintptr_t)tile_row_base<A, USES_CTBR>(S, tid, n);
  if (tid < 2 * kGroups) mbar_init(sbase + R::BAR_OFF + 8 * tid, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int stream = blockIdx.x * kGroups + g, n_streams = gridDim.x * kGroups;
  const int copy = (stream + gw) % DDQC_STAT_COPIES;
  const uint32_t bar_main = sbase + R::BAR_OFF + 16 * g;
  const uint32_t buf_addr = sbase + (uint32_t)(R::BUF_BYTES * g);
  const float* buf = reinterpret_cast<const float*>(tile_smem + R::BUF_BYTES * g);

  const int t0 = stream;
  pdl_launch_dependents();
  // The main rows were written by the previous STEP kernel, complete before the previous finalisation released this grid:
  // their copies start before the wait.  The late rows (the two episode sums a fix-up may touch, the PID state whose
  // need depends on the control block) start after it.
  if (gt == 0 && t0 < n_tiles) tile_issue_loads<A, USES_CTBR>(rowptr, actions, buf_addr, bar_main, t0, true, false, false);
  pdl_wait();
  const bool indep = (P.flags & DDQC_F_INDEPENDENT_ENVS) != 0u;
  const uint64_t ev = ctl->rng_event;
  const bool pid_pending = ctl->pid_reset_pending != 0u;
  const bool pid_rows = indep || !pid_pending;  // A17: a pending whole-batch reset makes the stored PID state moot
  if (gt == 0 && t0 < n_tiles) tile_issue_loads<A, USES_CTBR>(rowptr, actions, buf_addr, bar_main, t0, false, true, pid_rows);

  int it = 0;
  for (int tile = t0; tile < n_tiles; tile += n_streams, ++it) {
    float rew_total = 0.0f;
    bool did_timeout = false, below_ground = false;
    const int i = tile * kTile + gt;
    TileIO<A, USES_CTBR> io{buf, S, O, rowptr, actions, n, i, gt, g, buf_addr, bar_main, (uint32_t)it & 1u,
                            tile + n_streams < n_tiles ? tile + n_streams : -1, lane == 0 && gw == (it + 1) % kGroupWarps,
                            pid_rows, ev, pid_pending};
    env_step_body<ACT, true>(P, O, ctl, io, nullptr, n, i, true, lane, copy, rew_total, did_timeout, below_ground);
    warp_stats(P, ctl, lane, copy, rew_total, did_timeout, below_ground);
  }
}

// ---- the step kernel, one thread per env, state straight from / to global memory --------------------------
// LEAN = the throughput configuration: no Euler angles needed (thresholds >= 180 deg, yaw reward
// scale 0), no actuator / observation noise, no materialised reference buffers, no measurement
// override.  It is the same code with those branches compiled out (fewer registers, no spills).
#ifndef DDQC_ENV_LEAN_WARPS
#define DDQC_ENV_LEAN_WARPS 32  // resident warps per SM the lean instantiation is compiled for (32 -> 64 registers);
#endif                          // measured: 40 warps (48 registers, 96 B of spills) 72.7 us, 48 warps 84.7 us against 64.0 us
template <int ACT, bool LEAN, bool FUSED, int THREADS, bool TAIL>
// (the full-featured instantiation gets 80 registers instead of 64: it spilled ~250 bytes per thread at 64)
__global__ void __launch_bounds__(THREADS, (LEAN ? 32 * DDQC_ENV_LEAN_WARPS : 768) / THREADS)
env_step_kernel(const DdqcEnvParams P, const DdqcEnvState S, const DdqcEnvOut O, DdqcEnvCtl* __restrict__ ctl,
                const float* __restrict__ actions, const float* __restrict__ meas_override, const int n, const int i0) {
  constexpr int A = (ACT == DDQC_ACT_CTBR_FIXED_YAW) ? 3 : 4;
  constexpr int NOBS = 13 + A;
  constexpr bool USES_CTBR = (ACT != DDQC_ACT_ROTOR_RPMS);
  // Threads past the end (last CTA only) recompute env n-1 and skip every side effect, so the
  // body is straight-line code and the warp-level reductions below need no special casing.
  // (TAIL: the launch covers the ragged tail [i0, n) behind the tile kernel; a plain launch has i0 = 0 compiled in -
  // the extra live register made the 64-register instantiation spill.)
  const int gid = (TAIL ? i0 : 0) + blockIdx.x * THREADS + threadIdx.x;
  const bool live = gid < n;
  const int i = live ? gid : n - 1;
  const int lane = threadIdx.x & 31;
  const int copy = (blockIdx.x + (threadIdx.x >> 5)) % DDQC_STAT_COPIES;
  if (!FUSED) pdl_launch_dependents();  // the finalisation may be scheduled; it waits for this grid to complete

  float rew_total = 0.0f;
  bool did_timeout = false, below_ground = false;
  GlobalIO<A, FUSED> io{S, O, actions, n, i, live};
  env_step_body<ACT, LEAN>(P, O, ctl, io, meas_override, n, i, live, lane, copy, rew_total, did_timeout, below_ground);
  warp_stats(P, ctl, lane, copy, rew_total, did_timeout, below_ground);

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

// ---- the pair kernel: two envs per thread on packed fp32x2 arithmetic (ddqc_env_pair.cuh) -------------------------
// Lean configuration only.  The statement order is that of env_step_body; every value is the (env 2j, env 2j+1) pair.
#ifndef DDQC_ENV_PAIR_THREADS
#define DDQC_ENV_PAIR_THREADS 128
#endif
#ifndef DDQC_ENV_PAIR_BLOCKS
#define DDQC_ENV_PAIR_BLOCKS 4  // 128 registers per thread
#endif
constexpr int kPairThreads = DDQC_ENV_PAIR_THREADS;

__device__ __forceinline__ f2 ldp(const float* p) { return f2{__ldcg(reinterpret_cast<const unsigned long long*>(p))}; }
__device__ __forceinline__ void stp(float* p, f2 v) { *reinterpret_cast<unsigned long long*>(p) = v.v; }

template <int ACT, class IO>
__device__ __forceinline__ void env_step_pair_body(const DdqcEnvParams& P, const EnvParams2& PP, const DdqcEnvState& S,
                                                   const DdqcEnvOut& O, DdqcEnvCtl* __restrict__ ctl, IO& io, const int n,
                                                   const int i, const bool live, const int lane, const int copy) {
  constexpr int A = (ACT == DDQC_ACT_CTBR_FIXED_YAW) ? 3 : 4;
  constexpr int NOBS = 13 + A;
  constexpr bool USES_CTBR = (ACT != DDQC_ACT_ROTOR_RPMS);
  const PairMath M{PP};
  const bool indep = (P.flags & DDQC_F_INDEPENDENT_ENVS) != 0u;

  // ------------------------------------------------------------------ load state
  io.begin();
  V3p p{io.template ld<F_POS>(0), io.template ld<F_POS>(1), io.template ld<F_POS>(2)};
  Q4p q{io.template ld<F_QUAT>(0), io.template ld<F_QUAT>(1), io.template ld<F_QUAT>(2), io.template ld<F_QUAT>(3)};
  V3p v{io.template ld<F_VEL>(0), io.template ld<F_VEL>(1), io.template ld<F_VEL>(2)};
  V3p w{io.template ld<F_ANG>(0), io.template ld<F_ANG>(1), io.template ld<F_ANG>(2)};
  V3p cmd{io.template ld<F_CMD>(0), io.template ld<F_CMD>(1), io.template ld<F_CMD>(2)};
  f2 last_act[A], act[A];
#pragma unroll
  for (int j = 0; j < A; ++j) last_act[j] = io.template ld<F_LACT>(j);
  {
    float a0[A], a1[A];  // the two action rows are adjacent: 2A consecutive floats
    io.load_actions(a0, a1);
#pragma unroll
    for (int j = 0; j < A; ++j) act[j] = pk(clipf(a0[j], -P.clip_actions, P.clip_actions), clipf(a1[j], -P.clip_actions, P.clip_actions));  // A1
  }
  int2 ep_len = io.template ldi<F_EPLEN>();
  int2 hover = io.template ldi<F_HOVER>();
  f2 sums[DDQC_NUM_REWARDS];
#pragma unroll
  for (int k = 2; k < DDQC_NUM_REWARDS; ++k) sums[k] = io.template ld<F_SUMS>(k);

  // as in env_step_body: what the previous finalisation writes is read below the wait
  io.late_sync();
  const uint64_t ev = io.rng_event(ctl);
  const bool pid_pending = !indep && io.pid_reset_pending(ctl);
  sums[0] = io.template ld<F_SUMS>(0);
  sums[1] = io.template ld<F_SUMS>(1);
  f2 integ[3], prev[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) integ[j] = prev[j] = LIT(L_ZERO);
  if (USES_CTBR && !pid_pending) {  // A17
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      integ[j] = io.template ld<F_PIDI>(j);
      prev[j] = io.template ld<F_PIDE>(j);
    }
  }
  io.inputs_done();

  // ------------------------------------------------------------------ action -> rpm
  f2 rpm[4];
  {
    f2 y[A];
#pragma unroll
    for (int j = 0; j < A; ++j) {
      const f2 ex = (P.flags & DDQC_F_ACTION_LATENCY) ? last_act[j] : act[j];  // A2
      // A3: lo + ((x - (-1)) * span) / 2;  x - (-1) == x + 1 and / 2 == * 0.5 in IEEE arithmetic
      y[j] = add2(K2A(act_lo, j), M.mul2(M.mul2(add2(ex, LIT(L_ONE)), K2A(act_span, j)), LIT(L_HALF)));
    }
    if (USES_CTBR) {
      const V3p nq{neg2(q.x), neg2(q.y), neg2(q.z)};
      const V3p meas = M.rotate_by_quat(M.rotate_by_quat(w, q, nq), Q4p{q.w, nq.x, nq.y, nq.z}, V3p{q.x, q.y, q.z});  // A5
      f2 sp[3];
      sp[0] = y[1];
      sp[1] = y[2];
      sp[2] = (ACT == DDQC_ACT_CTBR) ? y[A - 1] : LIT(L_ZERO);  // A4
      const f2 ms[3] = {meas.x, meas.y, meas.z};
      f2 c[4];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const f2 err = sub2(sp[j], ms[j]);
        integ[j] = add2(integ[j], M.mul2(err, K2(step_dt)));  // A6
        const f2 deriv = M.mul2(sub2(err, prev[j]), K2(inv_step_dt));
        prev[j] = err;
        c[j] = add2(add2(M.mul2(K2A(pid_kp, j), err), M.mul2(K2A(pid_ki, j), integ[j])), M.mul2(K2A(pid_kd, j), deriv));
        if (live) {
          stp(S.pid_integral + j * n + i, integ[j]);
          stp(S.pid_prev_error + j * n + i, prev[j]);
        }
      }
      c[3] = y[0];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        f2 f = add2(add2(add2(M.mul2(K2A(mixer, 4 * r + 0), c[0]), M.mul2(K2A(mixer, 4 * r + 1), c[1])),
                         M.mul2(K2A(mixer, 4 * r + 2), c[2])),
                    M.mul2(K2A(mixer, 4 * r + 3), c[3]));
        f = pk(lo(f) < 0.0f ? 0.0f : lo(f), hi(f) < 0.0f ? 0.0f : hi(f));  // A7
        rpm[r] = M.mul2(sqrt2(f), K2(inv_sqrt_kf));
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) rpm[r] = y[r < A ? r : 0];
    }
  }

  // ------------------------------------------------------------------ physics (A8)
  const V3p last_pos = p;
  {
    // rotor_wrench
    const f2 s0 = M.mul2(rpm[0], rpm[0]), s1 = M.mul2(rpm[1], rpm[1]), s2 = M.mul2(rpm[2], rpm[2]), s3 = M.mul2(rpm[3], rpm[3]);
    const f2 t0 = M.mul2(s0, K2(kf)), t1 = M.mul2(s1, K2(kf)), t2 = M.mul2(s2, K2(kf)), t3 = M.mul2(s3, K2(kf));
    const f2 fz = add2(add2(add2(t0, t1), t2), t3);
    const f2 tau_x = fma2(t3, K2A(arm_x, 3), fma2(t2, K2A(arm_x, 2), fma2(t1, K2A(arm_x, 1), M.mul2(t0, K2A(arm_x, 0)))));
    const f2 tau_y = fma2(t3, K2A(arm_y, 3), fma2(t2, K2A(arm_y, 2), fma2(t1, K2A(arm_y, 1), M.mul2(t0, K2A(arm_y, 0)))));
    const f2 tau_z = fma2(s3, K2A(km_spin, 3), fma2(s2, K2A(km_spin, 2), fma2(s1, K2A(km_spin, 1), M.mul2(s0, K2A(km_spin, 0)))));
    const WrenchP W{M.mul2(K2A(k_w, 0), tau_x), M.mul2(K2A(k_w, 1), tau_y), M.mul2(K2A(k_w, 2), tau_z),
                    M.mul2(K2(two_dt_over_m), fz), M.mul2(K2(dt_over_m), fz)};
    for (int sidx = 0; sidx < P.decimation; ++sidx) {  // substep
      // pose with the old velocities, then the velocities with the forces at the new pose (ddqc_device.cuh::substep)
      p.x = fma2(v.x, K2(dt), p.x);
      p.y = fma2(v.y, K2(dt), p.y);
      p.z = fma2(v.z, K2(dt), p.z);
      const f2 hx = M.mul2(w.x, K2(half_dt)), hy = M.mul2(w.y, K2(half_dt)), hz = M.mul2(w.z, K2(half_dt));
      const f2 s = fma2(hz, hz, fma2(hy, hy, M.mul2(hx, hx)));
      f2 c = fma2(s, K2A(cos_coef, 4), K2A(cos_coef, 3));  // horner5
      c = fma2(c, s, K2A(cos_coef, 2));
      c = fma2(c, s, K2A(cos_coef, 1));
      c = fma2(c, s, K2A(cos_coef, 0));
      f2 k = fma2(s, K2A(sinc_coef, 4), K2A(sinc_coef, 3));
      k = fma2(k, s, K2A(sinc_coef, 2));
      k = fma2(k, s, K2A(sinc_coef, 1));
      k = fma2(k, s, K2A(sinc_coef, 0));
      const Q4p nq = M.quat_mul(q, V3p{neg2(q.x), neg2(q.y), neg2(q.z)}, Q4p{c, M.mul2(k, hx), M.mul2(k, hy), M.mul2(k, hz)});
      const f2 n2 = fma2(nq.z, nq.z, fma2(nq.y, nq.y, fma2(nq.x, nq.x, M.mul2(nq.w, nq.w))));
      const f2 inv_n = fma2(LIT(L_MINUS_HALF), n2, LIT(L_ONE_HALF));
      q = Q4p{M.mul2(nq.w, inv_n), M.mul2(nq.x, inv_n), M.mul2(nq.y, inv_n), M.mul2(nq.z, inv_n)};
      const f2 nwx = fma2(neg2(K2A(b_w, 0)), M.mul2(w.y, w.z), add2(w.x, W.ax));
      const f2 nwy = fma2(neg2(K2A(b_w, 1)), M.mul2(w.z, w.x), add2(w.y, W.ay));
      const f2 nwz = fma2(neg2(K2A(b_w, 2)), M.mul2(w.x, w.y), add2(w.z, W.az));
      w = V3p{nwx, nwy, nwz};
      v.x = fma2(W.cxy, fma2(q.x, q.z, M.mul2(q.w, q.y)), v.x);
      v.y = fma2(W.cxy, fma2(q.y, q.z, neg2(M.mul2(q.w, q.x))), v.y);
      const f2 r33 = fma2(LIT(L_MINUS_TWO), fma2(q.x, q.x, M.mul2(q.y, q.y)), LIT(L_ONE));
      v.z = fma2(W.cz, r33, sub2(v.z, K2(dt_g)));
    }
  }

  // ------------------------------------------------------------------ buffers (A9-A12)
  ep_len.x += 1;
  ep_len.y += 1;
  V3p rel{sub2(cmd.x, p.x), sub2(cmd.y, p.y), sub2(cmd.z, p.z)};
  V3p last_rel{sub2(cmd.x, last_pos.x), sub2(cmd.y, last_pos.y), sub2(cmd.z, last_pos.z)};
  const V3p nqv{neg2(q.x), neg2(q.y), neg2(q.z)}, qv{q.x, q.y, q.z};
  const Q4p qc{q.w, nqv.x, nqv.y, nqv.z};
  V3p lin_b = M.rotate_by_quat(v, qc, qv);
  V3p ang_b = M.rotate_by_quat(M.rotate_by_quat(w, q, nqv), qc, qv);

  // ------------------------------------------------------------------ hover / resample (A13-A15)
  b2 hover_resampled{false, false};
  V3p cmd_new = cmd;
  if (P.flags & DDQC_F_AUTO_TARGET) {
    const f2 dist = sqrt2(M.sumsq3(rel));
    const b2 at_target{lo(dist) < P.at_target_threshold, hi(dist) < P.at_target_threshold};
    b2 fire;
    if (P.min_hover_steps > 0) {
      hover.x = at_target.a ? hover.x + 1 : 0;
      hover.y = at_target.b ? hover.y + 1 : 0;
      fire = b2{hover.x >= P.min_hover_steps, hover.y >= P.min_hover_steps};
    } else {
      fire = at_target;
    }
    if (fire.a | fire.b) {
      V3 c0{lo(cmd.x), lo(cmd.y), lo(cmd.z)}, c1{hi(cmd.x), hi(cmd.y), hi(cmd.z)};
      if (fire.a) {
        c0 = resample_command(P, (uint32_t)i, ev, DDQC_STREAM_HOVER);
        hover.x = 0;
      }
      if (fire.b) {
        c1 = resample_command(P, (uint32_t)(i + 1), ev, DDQC_STREAM_HOVER);
        hover.y = 0;
      }
      cmd_new = V3p{pk(c0.x, c1.x), pk(c0.y, c1.y), pk(c0.z, c1.z)};
      hover_resampled = fire;
    }
  }

  // ------------------------------------------------------------------ termination (A8, A9); LEAN: no Euler thresholds
  auto crash_of = [&](auto part) {
    return (fabsf(part(rel.x)) > P.term_x) | (fabsf(part(rel.y)) > P.term_y) | (fabsf(part(rel.z)) > P.term_z) |
           (part(p.z) < P.term_ground) | (fabsf(part(ang_b.x)) > P.term_ang_vel) | (fabsf(part(ang_b.y)) > P.term_ang_vel) |
           (fabsf(part(ang_b.z)) > P.term_ang_vel) | (fabsf(part(lin_b.x)) > P.term_lin_vel) |
           (fabsf(part(lin_b.y)) > P.term_lin_vel) | (fabsf(part(lin_b.z)) > P.term_lin_vel);
  };
  // (the scalar kernel also ORs |euler| > threshold with euler = 0, which is false for every threshold >= 0 and for NaN)
  const b2 crash{(bool)crash_of([](f2 x) { return lo(x); }), (bool)crash_of([](f2 x) { return hi(x); })};
  const b2 below_ground{live && (lo(p.z) < 0.0f), live && (hi(p.z) < 0.0f)};
  const b2 did_timeout{live && (ep_len.x > P.max_episode_length), live && (ep_len.y > P.max_episode_length)};
  const b2 did_reset{live && (did_timeout.a || crash.a), live && (did_timeout.b || crash.b)};

  f2 last_act_obs[A];
#pragma unroll
  for (int j = 0; j < A; ++j) last_act_obs[j] = last_act[j];
  Q4p q_obs = q;

  // A18: sums of the episode sums over the envs that reset, one half of the pairs after the other
  {
    const unsigned int rm0 = __ballot_sync(0xffffffffu, did_reset.a), rm1 = __ballot_sync(0xffffffffu, did_reset.b);
    if (rm0 | rm1) {
      float part = 0.0f;
      for (unsigned int m = rm0; m; m &= m - 1) {
        const int L = __ffs(m) - 1;
#pragma unroll
        for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
          const float vk = __shfl_sync(0xffffffffu, lo(sums[k]), L);
          if (lane == k) part += vk;
        }
      }
      for (unsigned int m = rm1; m; m &= m - 1) {
        const int L = __ffs(m) - 1;
#pragma unroll
        for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
          const float vk = __shfl_sync(0xffffffffu, hi(sums[k]), L);
          if (lane == k) part += vk;
        }
      }
      if (lane < DDQC_NUM_REWARDS) atomicAdd(&ctl->acc_sums[copy][lane], part);
      if (lane == 8) atomicAdd(&ctl->acc_reset[copy], (unsigned int)(__popc(rm0) + __popc(rm1)));
    }
  }

  if (did_reset.a | did_reset.b) {  // reset_idx (hover_env.py:556-595), per lane of the pair
    const f2 zero = LIT(L_ZERO);
#pragma unroll
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) sums[k] = sel2(did_reset, zero, sums[k]);
    if (USES_CTBR && indep && live) {  // ctbr_controller.reset() of the env's own instance
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        stp(S.pid_integral + j * n + i, sel2(did_reset, zero, integ[j]));
        stp(S.pid_prev_error + j * n + i, sel2(did_reset, zero, prev[j]));
      }
    }
    const V3p ip{K2A(init_pos, 0), K2A(init_pos, 1), K2A(init_pos, 2)};
    p = V3p{sel2(did_reset, ip.x, p.x), sel2(did_reset, ip.y, p.y), sel2(did_reset, ip.z, p.z)};
    const V3p rel_r{sub2(cmd_new.x, ip.x), sub2(cmd_new.y, ip.y), sub2(cmd_new.z, ip.z)};  // A16, old command
    rel = V3p{sel2(did_reset, rel_r.x, rel.x), sel2(did_reset, rel_r.y, rel.y), sel2(did_reset, rel_r.z, rel.z)};
    last_rel = V3p{sel2(did_reset, rel_r.x, last_rel.x), sel2(did_reset, rel_r.y, last_rel.y), sel2(did_reset, rel_r.z, last_rel.z)};
    q = Q4p{sel2(did_reset, K2A(init_quat, 0), q.w), sel2(did_reset, K2A(init_quat, 1), q.x),
            sel2(did_reset, K2A(init_quat, 2), q.y), sel2(did_reset, K2A(init_quat, 3), q.z)};
    q_obs = q;
    v = V3p{sel2(did_reset, zero, v.x), sel2(did_reset, zero, v.y), sel2(did_reset, zero, v.z)};
    w = V3p{sel2(did_reset, zero, w.x), sel2(did_reset, zero, w.y), sel2(did_reset, zero, w.z)};
    lin_b = V3p{sel2(did_reset, zero, lin_b.x), sel2(did_reset, zero, lin_b.y), sel2(did_reset, zero, lin_b.z)};
    ang_b = V3p{sel2(did_reset, zero, ang_b.x), sel2(did_reset, zero, ang_b.y), sel2(did_reset, zero, ang_b.z)};
#pragma unroll
    for (int j = 0; j < A; ++j) last_act_obs[j] = sel2(did_reset, zero, last_act_obs[j]);
    V3 c0{lo(cmd_new.x), lo(cmd_new.y), lo(cmd_new.z)}, c1{hi(cmd_new.x), hi(cmd_new.y), hi(cmd_new.z)};
    if (did_reset.a) {
      ep_len.x = 0;
      hover.x = 0;
      c0 = resample_command(P, (uint32_t)i, ev, DDQC_STREAM_RESET);
      hover_resampled.a = false;
    }
    if (did_reset.b) {
      ep_len.y = 0;
      hover.y = 0;
      c1 = resample_command(P, (uint32_t)(i + 1), ev, DDQC_STREAM_RESET);
      hover_resampled.b = false;
    }
    cmd_new = V3p{pk(c0.x, c1.x), pk(c0.y, c1.y), pk(c0.z, c1.z)};
  }

  // ------------------------------------------------------------------ rewards (A19)
  const f2 r_target = sub2(M.sumsq3(last_rel), M.sumsq3(rel));
  f2 r_close = M.mul2(sub2(K2(at_target_threshold_sq), M.sumsq3(rel)), K2(inv_at_target_threshold_sq));
  r_close = pk(lo(r_close) < 0.0f ? 0.0f : lo(r_close), hi(r_close) < 0.0f ? 0.0f : hi(r_close));
  const f2 r_hover = pk((float)hover.x, (float)hover.y);
  f2 r_smooth;
  {
    const f2 d0 = sub2(act[0], last_act_obs[0]);
    r_smooth = M.mul2(d0, d0);
#pragma unroll
    for (int j = 1; j < A; ++j) {
      const f2 d = sub2(act[j], last_act_obs[j]);
      r_smooth = add2(r_smooth, M.mul2(d, d));
    }
  }
  const f2 r_yaw = LIT(L_ZERO);  // LEAN: yaw reward scale 0, no Euler angles
  const f2 r_ang = sqrt2(M.sumsq3(V3p{M.mul2(ang_b.x, K2(inv_pi_lit)), M.mul2(ang_b.y, K2(inv_pi_lit)), M.mul2(ang_b.z, K2(inv_pi_lit))}));
  const f2 r_crash = pk(crash.a ? 1.0f : 0.0f, crash.b ? 1.0f : 0.0f);
  const f2 rk[DDQC_NUM_REWARDS] = {M.mul2(r_target, K2A(rew_scale, 0)), M.mul2(r_close, K2A(rew_scale, 1)),
                                   M.mul2(r_hover, K2A(rew_scale, 2)),  M.mul2(r_smooth, K2A(rew_scale, 3)),
                                   M.mul2(r_yaw, K2A(rew_scale, 4)),    M.mul2(r_ang, K2A(rew_scale, 5)),
                                   M.mul2(r_crash, K2A(rew_scale, 6))};
  const f2 sum0_pre = sums[0], sum1_pre = sums[1];
  f2 rew = LIT(L_ZERO);
#pragma unroll
  for (int k = 0; k < DDQC_NUM_REWARDS; ++k) {
    rew = add2(rew, rk[k]);
    sums[k] = add2(sums[k], rk[k]);
  }
  const float rew_total = live ? lo(rew) + hi(rew) : 0.0f;

  // ------------------------------------------------------------------ observations (A20)
  f2 obs[NOBS];
  obs[0] = clip2(M.mul2(rel.x, K2(obs_scale_pos)), -1.0f, 1.0f);
  obs[1] = clip2(M.mul2(rel.y, K2(obs_scale_pos)), -1.0f, 1.0f);
  obs[2] = clip2(M.mul2(rel.z, K2(obs_scale_pos)), -1.0f, 1.0f);
  obs[3] = q_obs.w;
  obs[4] = q_obs.x;
  obs[5] = q_obs.y;
  obs[6] = q_obs.z;
  obs[7] = clip2(M.mul2(lin_b.x, K2(obs_scale_lin)), -1.0f, 1.0f);
  obs[8] = clip2(M.mul2(lin_b.y, K2(obs_scale_lin)), -1.0f, 1.0f);
  obs[9] = clip2(M.mul2(lin_b.z, K2(obs_scale_lin)), -1.0f, 1.0f);
  obs[10] = clip2(M.mul2(ang_b.x, K2(obs_scale_ang)), -1.0f, 1.0f);
  obs[11] = clip2(M.mul2(ang_b.y, K2(obs_scale_ang)), -1.0f, 1.0f);
  obs[12] = clip2(M.mul2(ang_b.z, K2(obs_scale_ang)), -1.0f, 1.0f);
#pragma unroll
  for (int j = 0; j < A; ++j) obs[13 + j] = last_act_obs[j];

  // A15/A16 fix-up records (rare): scalar arithmetic on the lane concerned, same expressions as env_step_body
  if ((hover_resampled.a | hover_resampled.b) && live) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!(h ? hover_resampled.b : hover_resampled.a)) continue;
      auto part = [&](f2 x) { return h ? hi(x) : lo(x); };
      const V3 rel_f{part(cmd_new.x) - part(p.x), part(cmd_new.y) - part(p.y), part(cmd_new.z) - part(p.z)};
      const V3 lrel_f{part(cmd_new.x) - part(last_pos.x), part(cmd_new.y) - part(last_pos.y), part(cmd_new.z) - part(last_pos.z)};
      const float rt = (sumsq3(lrel_f) - sumsq3(rel_f)) * P.rew_scale[0];
      float rc = (P.at_target_threshold_sq - sumsq3(rel_f)) * P.inv_at_target_threshold_sq;
      rc = (rc < 0.0f ? 0.0f : rc) * P.rew_scale[1];
      float rw = 0.0f + rt;
      rw = rw + rc;
#pragma unroll
      for (int k = 2; k < DDQC_NUM_REWARDS; ++k) rw = rw + part(rk[k]);
      const unsigned int slot = atomicAdd(&ctl->fixup_count, 1u);
      O.fixup_idx[slot] = i + h;
      float* fv = O.fixup_val + (size_t)slot * 8;
      fv[0] = rw;
      fv[1] = part(sum0_pre) + rt;
      fv[2] = part(sum1_pre) + rc;
      fv[3] = clipf(rel_f.x * P.obs_scale_pos, -1.0f, 1.0f);
      fv[4] = clipf(rel_f.y * P.obs_scale_pos, -1.0f, 1.0f);
      fv[5] = clipf(rel_f.z * P.obs_scale_pos, -1.0f, 1.0f);
      fv[6] = 0.0f;
      fv[7] = 0.0f;
    }
  }

  // ------------------------------------------------------------------ store
  if (live) {
    stp(S.pos + i, p.x);
    stp(S.pos + n + i, p.y);
    stp(S.pos + 2 * n + i, p.z);
    stp(S.quat + i, q.w);
    stp(S.quat + n + i, q.x);
    stp(S.quat + 2 * n + i, q.y);
    stp(S.quat + 3 * n + i, q.z);
    stp(S.vel + i, v.x);
    stp(S.vel + n + i, v.y);
    stp(S.vel + 2 * n + i, v.z);
    stp(S.ang + i, w.x);
    stp(S.ang + n + i, w.y);
    stp(S.ang + 2 * n + i, w.z);
    stp(S.commands + i, cmd_new.x);
    stp(S.commands + n + i, cmd_new.y);
    stp(S.commands + 2 * n + i, cmd_new.z);
#pragma unroll
    for (int j = 0; j < A; ++j) stp(S.last_actions + j * n + i, act[j]);  // hover_env.py:544, all envs
    *reinterpret_cast<int2*>(S.episode_length + i) = ep_len;
    *reinterpret_cast<int2*>(S.hover_counter + i) = hover;
#pragma unroll
    for (int k = 0; k < DDQC_NUM_REWARDS; ++k) stp(S.episode_sums + k * n + i, sums[k]);

    if (NOBS == 16) {
      float4* o4 = reinterpret_cast<float4*>(O.obs + (size_t)i * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) o4[c] = make_float4(lo(obs[4 * c]), lo(obs[4 * c + 1]), lo(obs[4 * c + 2]), lo(obs[4 * c + 3]));
#pragma unroll
      for (int c = 0; c < 4; ++c) o4[4 + c] = make_float4(hi(obs[4 * c]), hi(obs[4 * c + 1]), hi(obs[4 * c + 2]), hi(obs[4 * c + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < NOBS; ++j) {
        O.obs[(size_t)i * NOBS + j] = lo(obs[j]);
        O.obs[(size_t)(i + 1) * NOBS + j] = hi(obs[j]);
      }
    }
    stp(O.rew + i, rew);
    *reinterpret_cast<uchar2*>(O.reset + i) = make_uchar2(did_reset.a ? 1 : 0, did_reset.b ? 1 : 0);
    stp(O.time_outs + i, pk(did_timeout.a ? 1.0f : 0.0f, did_timeout.b ? 1.0f : 0.0f));
  }

  // ------------------------------------------------------------------ statistics (one RED per warp each)
  {
    const unsigned int tm = (unsigned int)(__popc(__ballot_sync(0xffffffffu, did_timeout.a)) + __popc(__ballot_sync(0xffffffffu, did_timeout.b)));
    float r = rew_total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 7) atomicAdd(&ctl->acc_sums[copy][7], r);
    if (lane == 9 && tm) atomicAdd(&ctl->acc_timeout[copy], tm);
    if (P.flags & DDQC_F_GROUND_WATCH) {
      const unsigned int gm = (unsigned int)(__popc(__ballot_sync(0xffffffffu, below_ground.a)) + __popc(__ballot_sync(0xffffffffu, below_ground.b)));
      if (lane == 10 && gm) atomicAdd(&ctl->acc_ground[copy], gm);
    }
  }
}


// the pair's state from global memory: one 64-bit access per row and thread
template <int A>
struct PairGlobalIO {
  const DdqcEnvState& S;
  const float* __restrict__ actions;
  const int n, i;
  template <int F>
  __device__ __forceinline__ const float* fptr() const {
    if constexpr (F == F_POS) return S.pos;
    else if constexpr (F == F_QUAT) return S.quat;
    else if constexpr (F == F_VEL) return S.vel;
    else if constexpr (F == F_ANG) return S.ang;
    else if constexpr (F == F_CMD) return S.commands;
    else if constexpr (F == F_LACT) return S.last_actions;
    else if constexpr (F == F_SUMS) return S.episode_sums;
    else if constexpr (F == F_PIDI) return S.pid_integral;
    else return S.pid_prev_error;
  }
  template <int F>
  __device__ __forceinline__ f2 ld(int c) const { return ldp(fptr<F>() + c * n + i); }
  template <int F>
  __device__ __forceinline__ int2 ldi() const {
    return __ldcg(reinterpret_cast<const int2*>((F == F_EPLEN ? S.episode_length : S.hover_counter) + i));
  }
  __device__ __forceinline__ void load_actions(float (&a0)[A], float (&a1)[A]) const {
    if (A == 3) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(actions + (size_t)i * 3);
      const f2 u0{src[0]}, u1{src[1]}, u2{src[2]};
      a0[0] = lo(u0), a0[1] = hi(u0), a0[2] = lo(u1), a1[0] = hi(u1), a1[1] = lo(u2), a1[2] = hi(u2);
    } else {
      const float4* src = reinterpret_cast<const float4*>(actions + (size_t)i * 4);
      const float4 r0 = src[0], r1 = src[1];
      a0[0] = r0.x, a0[1] = r0.y, a0[2] = r0.z, a0[A - 1] = r0.w;
      a1[0] = r1.x, a1[1] = r1.y, a1[2] = r1.z, a1[A - 1] = r1.w;
    }
  }
  __device__ __forceinline__ void begin() const {}
  __device__ __forceinline__ void late_sync() const { pdl_wait(); }
  __device__ __forceinline__ void inputs_done() const {}
  __device__ __forceinline__ uint64_t rng_event(const DdqcEnvCtl* ctl) const { return ctl->rng_event; }
  __device__ __forceinline__ bool pid_reset_pending(const DdqcEnvCtl* ctl) const { return ctl->pid_reset_pending != 0u; }
};

template <int ACT>
__global__ void __launch_bounds__(kPairThreads, DDQC_ENV_PAIR_BLOCKS)
env_step_pair_kernel(const DdqcEnvParams P, const EnvParams2 PP, const DdqcEnvState S, const DdqcEnvOut O,
                     DdqcEnvCtl* __restrict__ ctl, const float* __restrict__ actions, const int n) {
  constexpr int A = (ACT == DDQC_ACT_CTBR_FIXED_YAW) ? 3 : 4;
  const int n_pairs = n >> 1;
  const int gid = blockIdx.x * kPairThreads + threadIdx.x;
  const bool live = gid < n_pairs;
  const int i = 2 * (live ? gid : n_pairs - 1);  // first env of the pair; threads past the end redo the last pair, no side effects
  const int lane = threadIdx.x & 31;
  const int copy = (blockIdx.x + (threadIdx.x >> 5)) % DDQC_STAT_COPIES;
  pdl_launch_dependents();
  PairGlobalIO<A> io{S, actions, n, i};
  env_step_pair_body<ACT>(P, PP, S, O, ctl, io, n, i, live, lane, copy);
}

// the pair kernel on bulk-copy staged inputs: a group of kTile / 2 threads per kTile-env tile, kPairGroups groups per CTA
constexpr int kPairGroupThreads = kTile / 2;
constexpr int kPairGroups = 512 / kPairGroupThreads;  // 512 threads of 128 registers per SM
constexpr int kPairTileThreads = kPairGroupThreads * kPairGroups;

template <int A, bool USES_CTBR>
struct PairTileIO {
  using R = TileRows<A>;
  const float* buf;      // the group's buffer, as floats
  const unsigned long long* rowptr;
  const float* actions;
  const int gt, g;
  const uint32_t buf_addr, bar_main, parity;
  const int next_tile;   // < 0: this stream has no further tile
  const bool issuer;     // this thread issues the next tile's copies
  const bool pid_rows;
  const uint64_t ev;     // control-block values, read once per launch
  const bool pid_pending;
  template <int F>
  __device__ __forceinline__ f2 ld(int c) const {
    return f2{*reinterpret_cast<const unsigned long long*>(buf + (R::template row0<F>() + c) * kTile + 2 * gt)};
  }
  template <int F>
  __device__ __forceinline__ int2 ldi() const { return *reinterpret_cast<const int2*>(buf + R::template row0<F>() * kTile + 2 * gt); }
  __device__ __forceinline__ void load_actions(float (&a0)[A], float (&a1)[A]) const {
    const float* src = buf + R::ACT_OFF + 2 * gt * A;
#pragma unroll
    for (int j = 0; j < A; ++j) a0[j] = src[j], a1[j] = src[A + j];
  }
  __device__ __forceinline__ void begin() const { mbar_wait_parity(bar_main, parity); }
  __device__ __forceinline__ void late_sync() const { mbar_wait_parity(bar_main + 8, parity); }
  __device__ __forceinline__ void inputs_done() const {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kPairGroupThreads) : "memory");
    if (issuer && next_tile >= 0) tile_issue_loads<A, USES_CTBR>(rowptr, actions, buf_addr, bar_main, next_tile, true, true, pid_rows);
  }
  __device__ __forceinline__ uint64_t rng_event(const DdqcEnvCtl*) const { return ev; }
  __device__ __forceinline__ bool pid_reset_pending(const DdqcEnvCtl*) const { return pid_pending; }
};

template <int ACT>
__global__ void __launch_bounds__(kPairTileThreads, 1)
env_step_pair_tile_kernel(const DdqcEnvParams P, const EnvParams2 PP, const DdqcEnvState S, const DdqcEnvOut O,
                          DdqcEnvCtl* __restrict__ ctl, const float* __restrict__ actions, const int n, const int n_tiles) {
  constexpr int A = (ACT == DDQC_ACT_CTBR_FIXED_YAW) ? 3 : 4;
  constexpr bool USES_CTBR = (ACT != DDQC_ACT_ROTOR_RPMS);
  using R = TileRows<A>;
  static_assert(kPairGroups <= kGroups, "the shared-memory map of the tile kernel has room for kGroups buffers");
  extern __shared__ __align__(128) uint8_t tile_smem[];
  const int tid = threadIdx.x, g = tid / kPairGroupThreads, gt = tid % kPairGroupThreads, gw = gt >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(tile_smem);
  unsigned long long* rowptr = reinterpret_cast<unsigned long long*>(tile_smem + R::PTR_OFF);
  if (tid < R::NROWS) rowptr[tid] = (unsigned long long)(uintptr_t)tile_row_base<A, USES_CTBR>(S, tid, n);
  if (tid < 2 * kPairGroups) mbar_init(sbase + R::BAR_OFF + 8 * tid, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int stream = blockIdx.x * kPairGroups + g, n_streams = gridDim.x * kPairGroups;
  const int copy = (stream + gw) % DDQC_STAT_COPIES;
  const uint32_t bar_main = sbase + R::BAR_OFF + 16 * g;
  const uint32_t buf_addr = sbase + (uint32_t)(R::BUF_BYTES * g);
  const float* buf = reinterpret_cast<const float*>(tile_smem + R::BUF_BYTES * g);
  const int t0 = stream;
  pdl_launch_dependents();
  if (gt == 0 && t0 < n_tiles) tile_issue_loads<A, USES_CTBR>(rowptr, actions, buf_addr, bar_main, t0, true, false, false);
  pdl_wait();
  const bool indep = (P.flags & DDQC_F_INDEPENDENT_ENVS) != 0u;
  const uint64_t ev = ctl->rng_event;
  const bool pid_pending = ctl->pid_reset_pending != 0u;
  const bool pid_rows = indep || !pid_pending;
  if (gt == 0 && t0 < n_tiles) tile_issue_loads<A, USES_CTBR>(rowptr, actions, buf_addr, bar_main, t0, false, true, pid_rows);
  int it = 0;
  for (int tile = t0; tile < n_tiles; tile += n_streams, ++it) {
    const int i = tile * kTile + 2 * gt;
    PairTileIO<A, USES_CTBR> io{buf, rowptr, actions, gt, g, buf_addr, bar_main, (uint32_t)it & 1u,
                                tile + n_streams < n_tiles ? tile + n_streams : -1,
                                lane == 0 && gw == (it + 1) % (kPairGroupThreads / 32), pid_rows, ev, pid_pending};
    env_step_pair_body<ACT>(P, PP, S, O, ctl, io, n, i, true, lane, copy);
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
  auto launch_smem = [&](auto kernel, dim3 g, dim3 b, size_t smem, auto... args) {
    if (err != cudaSuccess) return;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = g;
    cfg.blockDim = b;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = (DDQC_ENV_FINALIZE == 2 && !fused) ? 1 : 0;
    err = cudaLaunchKernelEx(&cfg, kernel, args...);
  };
  auto launch = [&](auto kernel, dim3 g, dim3 b, auto... args) { launch_smem(kernel, g, b, (size_t)0, args...); };
  // tile kernel (bulk-copy staged, persistent): lean configuration, large batches, 16-byte aligned rows
  static const int64_t tile_min_n = [] {
    const char* e = getenv("DDQC_ENV_TILE_MIN_N");
    return e ? (int64_t)atoll(e) : (int64_t)DDQC_ENV_TILE_MIN_N;
  }();
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  static const int64_t pair_tile_min_n = [] {
    const char* e = getenv("DDQC_ENV_PAIR_TILE_MIN_N");
    return e ? (int64_t)atoll(e) : (int64_t)DDQC_ENV_PAIR_TILE_MIN_N;
  }();
  const bool tile_align = lean && !fused && (n % 4) == 0 && aligned16(S->pos) && aligned16(S->quat) &&
                       aligned16(S->vel) && aligned16(S->ang) && aligned16(S->commands) && aligned16(S->last_actions) &&
                       aligned16(S->episode_sums) && aligned16(S->episode_length) && aligned16(S->hover_counter) &&
                       aligned16(S->pid_integral) && aligned16(S->pid_prev_error) && aligned16(O->rew) &&
                       aligned16(O->time_outs) && aligned16(O->reset) && aligned16(actions);
  const bool pair_tile_ok = tile_align && n_env >= pair_tile_min_n;
  const bool tile_ok = tile_align && !pair_tile_ok && n_env >= tile_min_n;
  // pair kernel (two envs per thread, packed fp32x2): lean configuration, even batch size (8-byte aligned row pairs)
  static const int64_t pair_min_n = [] {
    const char* e = getenv("DDQC_ENV_PAIR_MIN_N");
    return e ? (int64_t)atoll(e) : (int64_t)DDQC_ENV_PAIR_MIN_N;
  }();
  auto aligned8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; };
  const bool pair_ok = lean && !fused && !tile_ok && !pair_tile_ok && n_env >= pair_min_n && (n % 2) == 0 && aligned8(S->pos) && aligned8(S->quat) &&
                       aligned8(S->vel) && aligned8(S->ang) && aligned8(S->commands) && aligned8(S->last_actions) &&
                       aligned8(S->episode_sums) && aligned8(S->episode_length) && aligned8(S->hover_counter) &&
                       aligned8(S->pid_integral) && aligned8(S->pid_prev_error) && aligned8(O->rew) && aligned8(O->time_outs) &&
                       (reinterpret_cast<uintptr_t>(O->reset) & 1u) == 0 && (reinterpret_cast<uintptr_t>(O->obs) & 15u) == 0 &&
                       (reinterpret_cast<uintptr_t>(actions) & 15u) == 0;
  EnvParams2 PP;
  if (pair_ok || pair_tile_ok) {
    const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(P);
    for (int k = 0; k < kEnvParamWords; ++k) PP.w[k] = (unsigned long long)wsrc[k] | ((unsigned long long)wsrc[k] << 32);
    const float lits[L_COUNT] = {-0.0f, 0.0f, 1.0f, 2.0f, 0.5f, 1.5f, -0.5f, -2.0f};
    for (int k = 0; k < L_COUNT; ++k) {
      uint32_t b;
      memcpy(&b, &lits[k], 4);
      PP.lit[k] = (unsigned long long)b | ((unsigned long long)b << 32);
    }
  }
  int sms = 148;
  if (tile_ok || pair_tile_ok) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
#define DDQC_LAUNCH_TILE(ACT)                                                                                  \
  do {                                                                                                        \
    using R = TileRows<(ACT == DDQC_ACT_CTBR_FIXED_YAW ? 3 : 4)>;                                             \
    static bool attr_set = false;                                                                             \
    if (!attr_set) {                                                                                          \
      err = cudaFuncSetAttribute(env_step_tile_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, R::SMEM_BYTES); \
      attr_set = err == cudaSuccess;                                                                          \
    }                                                                                                         \
    const int n_tiles = n / kTile;                                                                            \
    const int grid_t = (n_tiles + kGroups - 1) / kGroups < sms ? (n_tiles + kGroups - 1) / kGroups : sms;     \
    launch_smem(env_step_tile_kernel<ACT>, dim3(grid_t), dim3(kTileThreads), (size_t)R::SMEM_BYTES, *P, *S, *O, ctl, actions, n, \
                n_tiles);                                                                                     \
    if (n_tiles * kTile < n)                                                                                  \
      launch(env_step_kernel<ACT, true, false, kThreadsSplit, true>, dim3((n - n_tiles * kTile + kThreadsSplit - 1) / kThreadsSplit), \
             dim3(kThreadsSplit), *P, *S, *O, ctl, actions, meas_override, n, n_tiles * kTile);               \
    launch(env_finalize_kernel<(ACT == DDQC_ACT_CTBR_FIXED_YAW ? 16 : 17), true, (ACT != DDQC_ACT_ROTOR_RPMS)>, dim3(1), \
           dim3(32), *P, *S, *O, ctl, n);                                                                     \
  } while (0)
#define DDQC_LAUNCH2(ACT, LEANV)                                                                              \
  do {                                                                                                        \
    if (fused) {                                                                                              \
      launch(env_step_kernel<ACT, LEANV, true, kThreadsFused, false>, dim3((n + kThreadsFused - 1) / kThreadsFused),  \
             dim3(kThreadsFused), *P, *S, *O, ctl, actions, meas_override, n, 0);                             \
    } else {                                                                                                  \
      launch(env_step_kernel<ACT, LEANV, false, kThreadsSplit, false>, dim3((n + kThreadsSplit - 1) / kThreadsSplit), \
             dim3(kThreadsSplit), *P, *S, *O, ctl, actions, meas_override, n, 0);                             \
      launch(env_finalize_kernel<(ACT == DDQC_ACT_CTBR_FIXED_YAW ? 16 : 17), LEANV, (ACT != DDQC_ACT_ROTOR_RPMS)>, \
             dim3(1), dim3(32), *P, *S, *O, ctl, n);                                                          \
    }                                                                                                         \
  } while (0)
#define DDQC_LAUNCH_PAIR_TILE(ACT)                                                                             \
  do {                                                                                                        \
    using R = TileRows<(ACT == DDQC_ACT_CTBR_FIXED_YAW ? 3 : 4)>;                                             \
    static bool attr_set = false;                                                                             \
    if (!attr_set) {                                                                                          \
      err = cudaFuncSetAttribute(env_step_pair_tile_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, R::SMEM_BYTES); \
      attr_set = err == cudaSuccess;                                                                          \
    }                                                                                                         \
    const int n_tiles = n / kTile;                                                                            \
    const int grid_t = (n_tiles + kPairGroups - 1) / kPairGroups < sms ? (n_tiles + kPairGroups - 1) / kPairGroups : sms; \
    launch_smem(env_step_pair_tile_kernel<ACT>, dim3(grid_t), dim3(kPairTileThreads), (size_t)R::SMEM_BYTES, *P, PP, *S, *O, ctl, \
                actions, n, n_tiles);                                                                         \
    if (n_tiles * kTile < n)                                                                                  \
      launch(env_step_kernel<ACT, true, false, kThreadsSplit, true>, dim3((n - n_tiles * kTile + kThreadsSplit - 1) / kThreadsSplit), \
             dim3(kThreadsSplit), *P, *S, *O, ctl, actions, meas_override, n, n_tiles * kTile);               \
    launch(env_finalize_kernel<(ACT == DDQC_ACT_CTBR_FIXED_YAW ? 16 : 17), true, (ACT != DDQC_ACT_ROTOR_RPMS)>, dim3(1), \
           dim3(32), *P, *S, *O, ctl, n);                                                                     \
  } while (0)
#define DDQC_LAUNCH_PAIR(ACT)                                                                                  \
  do {                                                                                                        \
    launch(env_step_pair_kernel<ACT>, dim3((n / 2 + kPairThreads - 1) / kPairThreads), dim3(kPairThreads), *P, PP, *S, *O, ctl, \
           actions, n);                                                                                       \
    launch(env_finalize_kernel<(ACT == DDQC_ACT_CTBR_FIXED_YAW ? 16 : 17), true, (ACT != DDQC_ACT_ROTOR_RPMS)>, dim3(1), \
           dim3(32), *P, *S, *O, ctl, n);                                                                     \
  } while (0)
#define DDQC_LAUNCH(ACT)                                                                                      \
  do {                                                                                                        \
    if (pair_tile_ok)                                                                                         \
      DDQC_LAUNCH_PAIR_TILE(ACT);                                                                             \
    else if (tile_ok)                                                                                         \
      DDQC_LAUNCH_TILE(ACT);                                                                                  \
    else if (pair_ok)                                                                                         \
      DDQC_LAUNCH_PAIR(ACT);                                                                                  \
    else if (lean)                                                                                            \
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
#undef DDQC_LAUNCH_TILE
#undef DDQC_LAUNCH_PAIR
#undef DDQC_LAUNCH_PAIR_TILE
  return (int)err;
}
