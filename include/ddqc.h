/*
 * ddqc.h -- C ABI of libddqc.so, the B200 (sm_100a) hot path of
 * pavelacamposp/data-driven-quad-control.
 *
 * The reference has no FFI of its own: its boundary for this path is the Python
 * class surface (HoverEnv, DroneCTBRController, DroneTrackingController,
 * NonlinearDataDrivenMPCController).  The host-side mirrors of those classes in
 * `data_driven_quad_control_b200/` keep that surface and call the entry points
 * below through ctypes with `tensor.data_ptr()` / the current CUDA stream.
 * Each entry point cites the reference code it replaces (paths relative to
 * /root/reference/src/data_driven_quad_control/).
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in `_host`.
 *    PyTorch owns every buffer; the library never allocates or frees device
 *    memory and keeps no global state besides cached function attributes.
 *  - SoA field layout: a field with k components of n_env environments is
 *    stored as k contiguous rows of n_env floats (`field[c*n_env + i]`).
 *  - Return value: 0 = ok, <0 = bad argument (DDQC_E_*), >0 = cudaError_t of
 *    the launch.  No exceptions, no abort, no host synchronisation inside.
 *  - Fully asynchronous on `stream` (a cudaStream_t passed as void*).
 */
#ifndef DDQC_H
#define DDQC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDQC_ABI_VERSION 3

#define DDQC_E_NULL (-1)    /* required pointer is NULL */
#define DDQC_E_SHAPE (-2)   /* bad size / unsupported shape */
#define DDQC_E_ALIGN (-3)   /* pointer not aligned as required */
#define DDQC_E_CONFIG (-4)  /* unsupported configuration value */

/* EnvActionType (envs/hover_env_config.py:122-144) */
#define DDQC_ACT_ROTOR_RPMS 0
#define DDQC_ACT_CTBR 1
#define DDQC_ACT_CTBR_FIXED_YAW 2

/* DdqcEnvParams.flags */
#define DDQC_F_ACTION_LATENCY 1u    /* env_cfg["simulate_action_latency"]  (hover_env.py:386-388) */
#define DDQC_F_AUTO_TARGET 2u       /* auto_target_updates                 (hover_env.py:473-478) */
#define DDQC_F_NEED_EULER 4u        /* Euler angles feed a live threshold or a non-zero yaw reward */
#define DDQC_F_GROUND_WATCH 16u     /* count the env-steps that end BELOW the ground plane z = 0.  The reference's scene has a
                                     * plane entity with collision on (hover_env.py:199); this restatement has no contact
                                     * model, so when a configuration disables the low-altitude termination
                                     * (termination_if_close_to_ground <= 0: the tracking example, the DD-MPC evaluation and
                                     * grid-search scripts) a drone that the reference would leave resting on the floor falls
                                     * through it here.  The count (DdqcEnvCtl.total_below_ground) lets callers detect that a
                                     * run has left the regime in which the two agree instead of differing silently */
#define DDQC_F_INDEPENDENT_ENVS 8u  /* every env behaves like its own num_envs = 1 HoverEnv instance: a reset zeroes the rate
                                     * PID of that env only and refreshes rel_pos of that env only.  The reference's two
                                     * batch-global reset side effects (hover_env.py:563-564, 592-593) couple the drones of one
                                     * instance; its grid search steps one instance with num_envs = num_processes drones
                                     * (dd_mpc_param_grid_search.py:282-291), so concurrent evaluation runs perturb each other
                                     * in a scheduling-dependent way.  A batch of thousands of runs in one instance would
                                     * couple them all; with this flag a run is what it is with num_processes = 1 */

#define DDQC_NUM_REWARDS 7          /* target, closeness, hover_time, smooth, yaw, angular, crash */
#define DDQC_STAT_COPIES 32         /* spread of the per-block statistics atomics */
#define DDQC_EP_RING 1024           /* rows of the extras["episode"] ring buffer */

/* Philox-4x32-10 stream ids (counter word 3); counter = (env, event_lo, event_hi, stream | sub<<8) */
#define DDQC_STREAM_HOVER 0u
#define DDQC_STREAM_RESET 1u
#define DDQC_STREAM_ACT_NOISE 2u
#define DDQC_STREAM_OBS_NOISE 3u

/* Host-side parameter block, passed by pointer and copied into the launch.
 * Every float is the fp32 rounding of the double the reference computes on the host. */
typedef struct DdqcEnvParams {
  int32_t action_type;          /* DDQC_ACT_* */
  int32_t decimation;           /* env_cfg["decimation"]                      hover_env.py:88,437 */
  int32_t max_episode_length;   /* ceil(episode_length_s / step_dt)           hover_env.py:156-158 */
  int32_t min_hover_steps;      /* ceil(min_hover_time_s / step_dt)           hover_env.py:303-304 */
  uint32_t flags;               /* DDQC_F_* */
  float clip_actions;           /* hover_env.py:379-383 */
  float act_lo[4];              /* action_bounds[:,0]                         hover_env.py:119-151 */
  float act_span[4];            /* fp32(action_bounds[:,1] - action_bounds[:,0]) */
  /* CTBR controller (controllers/ctbr/ctbr_controller.py:221-264) */
  float step_dt;                /* PID dt = dt*decimation                     hover_env.py:105-111 */
  float inv_step_dt;            /* 1.0f / step_dt  (torch CUDA tensor/scalar division semantics) */
  float pid_kp[3], pid_ki[3], pid_kd[3];
  float mixer[16];              /* row-major pinv(A) @ diag(J,1), computed on the host with torch */
  float inv_sqrt_kf;
  /* rigid body (restated Genesis substep; see DESIGN.md and oracle/drone_physics.py) */
  float dt, half_dt;            /* sim dt, dt/2 */
  float dt_g;                   /* dt * 9.81 */
  float kf;                     /* thrust coefficient, N / RPM^2 */
  float dt_over_m, two_dt_over_m;
  float k_w[3];                 /* dt / J_ii */
  float b_w[3];                 /* dt (J_kk - J_jj) / J_ii, (i,j,k) cyclic */
  float arm_x[4], arm_y[4];     /* roll / pitch torque arm of rotor i:  y_i, -x_i */
  float km_spin[4];             /* KM * spin_i */
  float cos_coef[5], sinc_coef[5]; /* Taylor coefficients of cos(a), sin(a)/a in a^2 */
  /* noise (hover_env.py:430-434, 610-621) */
  float actuator_noise_std, min_rpm, max_rpm, obs_noise_std;
  /* termination thresholds (hover_env.py:481-520) */
  float term_pitch, term_roll, term_x, term_y, term_z, term_ground, term_ang_vel, term_lin_vel;
  float init_pos[3], init_quat[4], inv_init_quat[4];
  float at_target_threshold, at_target_threshold_sq, inv_at_target_threshold_sq;
  float inv_pi_lit, inv_180;    /* 1.0f/3.14159f, 1.0f/180.0f             hover_env.py:847,852 */
  float cmd_lo[3], cmd_span[3]; /* command_cfg ranges: lower, fp32(upper-lower)   hover_env.py:328-337 */
  float obs_scale_pos, obs_scale_lin, obs_scale_ang;
  float rew_scale[DDQC_NUM_REWARDS]; /* reward_scales[k]*step_dt, dict order   hover_env.py:246-247 */
  float yaw_lambda;
  float episode_length_s;
  uint64_t seed;                /* Philox key */
} DdqcEnvParams;

/* Device control block (one per env instance; zero-initialised by the host). */
typedef struct DdqcEnvCtl {
  uint64_t rng_event;           /* bumped by one per step / external reset launch */
  uint64_t step_count;
  uint64_t total_resets, total_timeouts, total_env_steps;
  double total_reward;
  double total_episode_sums[DDQC_NUM_REWARDS]; /* sum of episode_sums[k] over all finished episodes */
  uint64_t total_below_ground;  /* env-steps that ended with z < 0 (DDQC_F_GROUND_WATCH) */
  uint64_t reserved0;
  uint32_t blocks_done;
  uint32_t pid_reset_pending;   /* A17: >=1 env reset in the previous launch -> whole-batch PID zero */
  uint32_t n_reset_last;
  uint32_t fixup_count;
  uint32_t ep_row;              /* row of ep_ring written by the last launch (= its DdqcEnvOut.ep_row) */
  uint32_t pad0;
  uint32_t acc_reset[DDQC_STAT_COPIES];
  uint32_t acc_timeout[DDQC_STAT_COPIES];
  uint32_t acc_ground[DDQC_STAT_COPIES];
  float acc_sums[DDQC_STAT_COPIES][8]; /* 7 episode sums of reset envs + sum of rewards */
  float ep_ring[DDQC_EP_RING][8];      /* mean(episode_sums[k][reset]) / episode_length_s per launch */
} DdqcEnvCtl;

/* Persistent per-env state, SoA. */
typedef struct DdqcEnvState {
  float* pos;             /* [3][n]  world position          (Genesis qpos[0:3]) */
  float* quat;            /* [4][n]  w,x,y,z                 (Genesis qpos[3:7]) */
  float* vel;             /* [3][n]  world linear velocity   (Genesis dofs 0..2) */
  float* ang;             /* [3][n]  body angular velocity   (Genesis dofs 3..5) */
  float* commands;        /* [3][n]  target position         hover_env.py:268-272 */
  float* last_actions;    /* [A][n]                          hover_env.py:279, 544 */
  float* pid_integral;    /* [3][n]  (CTBR modes)            vectorized_pid_controller.py:86-95 */
  float* pid_prev_error;  /* [3][n] */
  float* episode_sums;    /* [7][n]                          hover_env.py:245-251 */
  int32_t* episode_length;/* [n]                             hover_env.py:265-267 */
  int32_t* hover_counter; /* [n]                             hover_env.py:309-311 */
} DdqcEnvState;

/* Outputs of one step.  Pointers in the second group may be NULL (not written). */
typedef struct DdqcEnvOut {
  float* obs;             /* [n][num_obs] row-major          hover_env.py:597-627 */
  float* rew;             /* [n]                             hover_env.py:535-539 */
  uint8_t* reset;         /* [n] 0/1                         hover_env.py:521-523 */
  float* time_outs;       /* [n] 0.0/1.0                     hover_env.py:525-530 */
  /* optional materialised reference buffers (SoA) */
  float* base_lin_vel;    /* [3][n] body frame               hover_env.py:464-467 */
  float* base_ang_vel;    /* [3][n] body frame               hover_env.py:468-470 */
  float* rel_pos;         /* [3][n]                          hover_env.py:453, 563 */
  float* last_rel_pos;    /* [3][n]                          hover_env.py:454, 564 */
  float* base_euler;      /* [3][n] degrees                  hover_env.py:456-463 */
  uint8_t* crash;         /* [n]                             hover_env.py:481-520 */
  float* rpm;             /* [4][n] rotor RPM applied this step */
  float* last_base_pos;   /* [3][n]                          hover_env.py:451, 562 */
  /* strict-reference fix-up list (A15/A16), capacity n */
  int32_t* fixup_idx;     /* [n] */
  float* fixup_val;       /* [n][8]: rew, sum_target, sum_closeness, obs0..2, rel(aux) unused */
  /* Row of ctl->ep_ring that receives this launch's extras["episode"] values (hover_env.py:583-589), taken modulo
   * DDQC_EP_RING.  The HOST names the row (step / reset_idx / restore_state each use the next one), so the host's
   * views of the ring can never drift from the device: a replayed CUDA graph rewrites exactly the rows whose views
   * were handed out during capture.  "No reset in this launch" copies the values of ctl->ep_row, the row the
   * previous launch wrote (the reference leaves the previous dict in place). */
  uint32_t ep_row;
  uint32_t pad0;
} DdqcEnvOut;

/* ---------------------------------------------------------------------------------
 * Vectorized environment (envs/hover_env.py)
 * --------------------------------------------------------------------------------- */

/* HoverEnv.step (hover_env.py:376-547): action clip/latency/decode, CTBR rate PID + mixer,
 * `decimation` rigid-body substeps, buffer refresh, hover counter + target resample,
 * termination, auto-reset (reset_idx, hover_env.py:556-595), rewards, observations,
 * episode statistics.  `actions` is [n][A] row-major.  `ang_vel_meas_override`
 * ([3][n] SoA, or NULL) replaces the body-rate measurement recomputed from the state for
 * this launch only (needed right after restore_from_state, hover_env.py:729-735). */
int ddqc_env_step(const DdqcEnvParams* params_host, const DdqcEnvState* state_host,
                  const DdqcEnvOut* out_host, DdqcEnvCtl* ctl, const float* actions,
                  const float* ang_vel_meas_override, int64_t n_env, void* stream);

/* HoverEnv.reset_idx called from outside step (hover_env.py:371-374, 556-595).
 * `idx` is [n_idx] int64 (unique indices).  Resets pose/velocities/buffers of idx, zeroes
 * the whole-batch PID state, resamples commands of idx, and writes
 * mean(episode_sums[k][idx])/episode_length_s to the ctl ring. */
int ddqc_env_reset_idx(const DdqcEnvParams* params_host, const DdqcEnvState* state_host,
                       const DdqcEnvOut* out_host, DdqcEnvCtl* ctl, const int64_t* idx,
                       int64_t n_idx, int64_t n_env, void* stream);

/* HoverEnv.get_current_state (hover_env.py:680-695): gathers idx into AoS row-major outputs
 * base_pos[n_idx][3], base_quat[n_idx][4], base_lin_vel[n_idx][3], base_ang_vel[n_idx][3] (body
 * frame), commands[n_idx][3], episode_length[n_idx], last_actions[n_idx][A]. */
int ddqc_env_get_state(const DdqcEnvParams* params_host, const DdqcEnvState* state_host,
                       const int64_t* idx, int64_t n_idx, int64_t n_env, float* base_pos,
                       float* base_quat, float* base_lin_vel, float* base_ang_vel,
                       float* commands, int32_t* episode_length, float* last_actions,
                       const float* ang_vel_meas_override, const float* lin_vel_override,
                       void* stream);

/* HoverEnv.restore_from_state (hover_env.py:697-753): scatters the AoS inputs to idx.  The
 * body-frame velocities are written into the simulator dofs as the reference does
 * (hover_env.py:729-735).  `commands` only enters the rel_pos / last_rel_pos buffers: the reference never
 * writes self.commands here (hover_env.py:712-716), the env keeps its current target.  Episode sums of idx
 * are averaged into the ctl ring and zeroed. */
int ddqc_env_restore_state(const DdqcEnvParams* params_host, const DdqcEnvState* state_host,
                           const DdqcEnvOut* out_host, DdqcEnvCtl* ctl, const int64_t* idx,
                           int64_t n_idx, int64_t n_env, const float* base_pos,
                           const float* base_quat, const float* base_lin_vel,
                           const float* base_ang_vel, const float* commands,
                           const int32_t* episode_length, const float* last_actions,
                           void* stream);

/* HoverEnv.update_target_pos (hover_env.py:667-677). target_pos is [n_idx][3] or, when
 * `broadcast` != 0, a single [3] row applied to every idx. */
int ddqc_env_update_target(const DdqcEnvState* state_host, const int64_t* idx, int64_t n_idx,
                           int64_t n_env, const float* target_pos, int broadcast, void* stream);

/* Body-frame velocities of every env from the state (hover_env.py:464-470), SoA [3][n]. */
int ddqc_env_body_vel(const DdqcEnvState* state_host, int64_t n_env, float* base_lin_vel,
                      float* base_ang_vel, void* stream);

/* Applies a pending whole-batch PID reset (A17) to the stored PID state and clears the flag,
 * so that host code reading the controller state sees what the reference holds. */
int ddqc_env_flush_pid_reset(const DdqcEnvState* state_host, DdqcEnvCtl* ctl, int64_t n_env,
                             void* stream);

/* ---------------------------------------------------------------------------------
 * Stand-alone controllers (AoS tensors with explicit element strides so that
 * transposed views need no copy)
 * --------------------------------------------------------------------------------- */

/* VectorizedPIDController.compute (utilities/vectorized_pid_controller.py:97-126) for
 * k <= 4 controllers per env: out = kp*e + ki*I + kd*d; integral/prev_error updated in place. */
int ddqc_pid_compute(const float* setpoints, int64_t sp_rs, int64_t sp_cs,
                     const float* measurements, int64_t ms_rs, int64_t ms_cs,
                     float* integral, float* prev_error, /* [n][k] contiguous */
                     const float* gains_host /* [k][3] kp,ki,kd */, float dt, float* out /* [n][k] */,
                     int64_t n, int32_t k, void* stream);

/* DroneCTBRController.compute (controllers/ctbr/ctbr_controller.py:221-264): rate PID ->
 * mixer -> clamp>=0 -> rpm = sqrt(F)*rsqrt(KF).  rpm is [n][4] row-major. */
int ddqc_ctbr_compute(const float* rate_meas, int64_t m_rs, int64_t m_cs,
                      const float* rate_sp, int64_t s_rs, int64_t s_cs,
                      const float* thrust_sp, int64_t t_s,
                      float* integral, float* prev_error, /* [n][3] contiguous */
                      const float* mixer_host /* [16] */, const float* gains_host /* [3][3] */,
                      float dt, float inv_sqrt_kf, float* rpm, int64_t n, void* stream);

/* DroneTrackingController.compute (controllers/tracking/tracking_controller.py:90-164):
 * position PID -> gravity compensation -> thrust along R(q) -> desired attitude ->
 * e_R = vee(0.5 (Rd^T R - R^T Rd)) -> body rates.  out is [n][4] = [thrust, wx, wy, wz]. */
int ddqc_tracking_compute(const float* pos, int64_t p_rs, int64_t p_cs,
                          const float* quat, int64_t q_rs, int64_t q_cs,
                          const float* pos_sp, int64_t ps_rs, int64_t ps_cs,
                          const float* quat_sp, int64_t qs_rs, int64_t qs_cs,
                          float* integral, float* prev_error, /* [n][3] contiguous */
                          const float* gains_host /* [3][3] */, float dt, float mass, float kR,
                          float* out, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------
 * Batched nonlinear data-driven MPC (third-party direct_data_driven_mpc
 * NonlinearDataDrivenMPCController; contract from the reference call sites
 * data_driven_mpc/utilities/param_grid_search/controller_creation.py:91-112 and
 * controller_evaluation.py:345-450)
 * --------------------------------------------------------------------------------- */

typedef struct DdqcMpcConfig {
  int32_t n, m, p, L, N;        /* estimated order, inputs, outputs, horizon, window length */
  int32_t alpha_reg_type;       /* 0 approximated, 1 previous, 2 zero */
  int32_t max_iter;             /* pivoting passes before the interior-point fallback (<=0: 12) */
  int32_t gram_refresh;         /* Gram cache: incremental updates between regenerations (<=0: 128) */
  int32_t alpha_sr_anchor;      /* alpha_reg_type 0 only - which equilibrium alpha_sr belongs to (the third-party source is not
                                 * available, the published scheme leaves it to the implementation):
                                 * 0 = the optimal reachable equilibrium for y_r (6-variable box QP; DESIGN.md, default),
                                 * 1 = the optimal equilibrium (u_s, y_s) of the PREVIOUS solve, read from us_opt / ys_opt as the
                                 *     previous launch left them (alpha_sr = 0 on a controller's first solve, have_prev == 0) */
  int32_t reserved0;
  double Q[8], R[8], S[8];      /* diagonal weights (first p / m / p entries) */
  double lamb_alpha, lamb_sigma, lamb_alpha_s, lamb_sigma_s;
  double U_lo[8], U_hi[8], Us_lo[8], Us_hi[8];
} DdqcMpcConfig;

/* Bytes of per-batch device workspace `ws` needed by ddqc_ddmpc_solve for `batch` controllers. */
int64_t ddqc_ddmpc_workspace_bytes(const DdqcMpcConfig* cfg_host, int64_t batch);

/* update_and_solve_data_driven_mpc for a batch of controllers.
 *  u_win [B][N][m], y_win [B][N][p] : rolling input/output windows (float64)
 *  y_r   [B][p]                      : output setpoints
 *  have_prev [B] int32               : 0 -> cold active set (first solve); alpha_sr (alpha_reg_type 0) is the
 *                                      alpha of the optimal reachable equilibrium for y_r and needs no history;
 *                                      < 0 -> the controller is skipped, its outputs, status, active set and Gram
 *                                      cache stay untouched (an evaluation run the reference would have aborted,
 *                                      controller_evaluation.py:424-447)
 *  lamb [B][4] (or NULL)            : per-controller (lamb_alpha, lamb_sigma, lamb_alpha_s, lamb_sigma_s)
 *                                      overriding the config values (grid search over the regularisation
 *                                      weights, configs/data_driven_mpc/dd_mpc_grid_search_params.yaml:92-111)
 *  act_state [B][nx] int8 (or NULL)  : active set of the box QP (-1 lower, 0 free, +1 upper),
 *                                      nx = (L-n) m + m + p; read as the warm start (shifted one
 *                                      step) when have_prev != 0, always written
 *  u_opt [B][L+1][m]                 : optimal predicted inputs ubar_0..ubar_L
 *  us_opt [B][m], ys_opt [B][p]      : optimal equilibrium
 *  status [B] int32 (0 ok, 1 not converged, 2 factorisation breakdown)
 *  iters [B] int32 = pivoting passes + (interior-point iterations << 16)
 *  ws: ddqc_ddmpc_workspace_bytes() bytes, 256-byte aligned; its first 256 bytes are the work counter */
int ddqc_ddmpc_solve(const DdqcMpcConfig* cfg_host, void* ws, const double* u_win,
                     const double* y_win, const double* y_r, const int32_t* have_prev, const double* lamb,
                     int8_t* act_state,
                     double* u_opt, double* us_opt, double* ys_opt, int32_t* status, int32_t* iters,
                     int64_t batch, void* gram_cache, void* stream);

/* Optional inputs / outputs of a solve for the modes every shipped reference config leaves off
 * (configs/data_driven_mpc/dd_mpc_controller_params.yaml:60-70 `alpha_reg_type: 1`, :85-90 `update_cost_threshold`;
 * ctor keywords data_driven_mpc/utilities/param_grid_search/controller_creation.py:108-110).  All members may be NULL.
 *  u_ic [B][N][m], y_ic [B][N][p] : windows whose LAST n samples are the initial condition, when u_win / y_win hold a
 *                                   frozen snapshot of the Hankel data (update_cost_threshold); not with a Gram cache
 *  v_sr [B][(m+p) ell + 1]        : alpha_reg_type 1 (PREVIOUS), required there: H alpha_sr, Hankel row (signal a,
 *                                   depth j) at a ell + j (inputs first), ones row last - ddqc_ddmpc_hankel_apply
 *  sr_out [B][m + p + p ell]      : alpha_reg_type 0: (u_sr, y_sr, e) of the alpha_sr sub-problem, e by (depth j,
 *                                   output c) at j p + c; v_sr = rho - e.  Input of ddqc_ddmpc_recover. */
typedef struct DdqcMpcAux {
  const double* u_ic;
  const double* y_ic;
  const double* v_sr;
  double* sr_out;
} DdqcMpcAux;

/* ddqc_ddmpc_solve with the optional block above (aux NULL = ddqc_ddmpc_solve). */
int ddqc_ddmpc_solve_ex(const DdqcMpcConfig* cfg_host, void* ws, const double* u_win,
                        const double* y_win, const double* y_r, const int32_t* have_prev, const double* lamb,
                        int8_t* act_state, double* u_opt, double* us_opt, double* ys_opt, int32_t* status,
                        int32_t* iters, int64_t batch, void* gram_cache, const DdqcMpcAux* aux_host, void* stream);

/* v[b] = H_b alpha[b] for the Hankel matrix of (u_win, y_win): alpha [B][C], C = N - ell + 1; v in the layout of
 * DdqcMpcAux.v_sr.  alpha_reg_type 1: alpha_sr = the previous optimal alpha (zeros before the first solve). */
int ddqc_ddmpc_hankel_apply(const DdqcMpcConfig* cfg_host, const double* u_win, const double* y_win,
                            const double* alpha, double* v, int64_t batch, void* stream);

/* Dual recovery after a solve (the condensed solver never forms alpha; the two optional modes need it).  With the
 * optimum (u_opt, us_opt, ys_opt) every Hankel row has a prescribed value (inputs, ones row: D = 0) or a target with
 * weight w (outputs: D = lamb_alpha / w); alpha* = alpha_sr + H'z with (G + D) z = t - v_sr (one Cholesky per
 * controller, one CTA each; numpy mirror oracle/ddmpc_condensed.py::recover_dual).
 *  u_win, y_win, u_ic, y_ic, y_r, have_prev (< 0: skipped), lamb, u_opt, us_opt, ys_opt : as given to / left by the solve
 *  v_sr (alpha_reg_type 1) / sr (alpha_reg_type 0: the solve's sr_out) / neither (type 2) : what alpha_sr was
 *  alpha_io [B][C] or NULL   : += H'z  (holds alpha_sr on entry, alpha* on return - alpha_reg_type 1)
 *  cost_out [B] or NULL      : the tracking cost  sum_k |ubar_k - u_s|_R^2 + |ybar_k - y_s|_Q^2 + |y_s - y_r|_S^2
 *                              (k = -n..L) that update_cost_threshold compares; +inf if the factorisation broke down
 *  ws: ddqc_ddmpc_recover_workspace_bytes() bytes, 256-byte aligned. */
int64_t ddqc_ddmpc_recover_workspace_bytes(const DdqcMpcConfig* cfg_host, int64_t batch);
int ddqc_ddmpc_recover(const DdqcMpcConfig* cfg_host, void* ws, const double* u_win, const double* y_win,
                       const double* u_ic, const double* y_ic, const double* y_r, const int32_t* have_prev,
                       const double* lamb, const double* u_opt, const double* us_opt, const double* ys_opt,
                       const double* v_sr, const double* sr, double* alpha_io, double* cost_out, int64_t batch,
                       void* stream);

/* update_cost_threshold: before a solve, copy the rolling windows into the Hankel-data snapshot of every controller
 * whose last tracking cost exceeds `threshold` or that has not been solved yet (have_prev == 0); the others keep
 * their data ("online input-output data updates are disabled when the tracking cost value is less than this value",
 * dd_mpc_controller_params.yaml:85-90).  refreshed [B] int32 (or NULL) reports which were copied. */
int ddqc_ddmpc_refresh_data(const DdqcMpcConfig* cfg_host, const double* u_roll, const double* y_roll,
                            double* u_data, double* y_data, const double* cost, double threshold,
                            const int32_t* have_prev, int32_t* refreshed, int64_t batch, void* stream);

/* Optional per-controller Gram cache (`gram_cache`, or NULL, of ddqc_ddmpc_solve / _store / _post_step): the block-packed
 * Gram matrix of every controller's Hankel matrix persists between solves; a solve whose window moved by exactly one
 * sample (one store_input_output_measurement) since the cached matrix applies the rank-2 correction
 * G' = G - (old first column)(..)' + (new last column)(..)' instead of regenerating G from the window.  The buffer
 * (ddqc_ddmpc_gram_cache_bytes() bytes, 16-byte aligned, ZEROED by the caller before the first solve and whenever the
 * windows are rewritten other than through _store / _post_step) belongs to the caller like every other buffer. */
int64_t ddqc_ddmpc_gram_cache_bytes(const DdqcMpcConfig* cfg_host, int64_t batch);

/* store_input_output_measurement for a batch: shift each window one sample and append. */
int ddqc_ddmpc_store(const DdqcMpcConfig* cfg_host, double* u_win, double* y_win,
                     const double* u_new, const double* y_new, int64_t batch, void* gram_cache, void* stream);

/* Closed-loop evaluation of a batch of controllers without host round trips
 * (sim_nonlinear_dd_mpc_control_loop_parallel, data_driven_mpc/utilities/param_grid_search/
 * controller_evaluation.py:345-450):
 *  ddqc_ddmpc_action    : env action in [-1,1] (fp32, [B][m]) from ubar_{n_step} via
 *                         linear_interpolate (utilities/math_utils.py:13-20) over `bounds` [m][2];
 *                         also copies the applied input (physical units, float64) to u_applied [B][m]
 *  ddqc_ddmpc_post_step : store_input_output_measurement(u_applied, y_meas) (:416-422) fused with the
 *                         tracking-error bookkeeping: sq_sum[b] += |y - y_r|^2, n_acc[b] += 1 and
 *                         fail_step[b] = step when |y - y_r| - d0[b] > max_inc (:424-447; max_inc < 0
 *                         disables the check).  Controllers with fail_step[b] >= 0 are frozen.
 *                         y_meas is fp32 with strides (y_rs, y_cs) in elements; y_log (or NULL) is
 *                         [steps][B][p] float64.  RMSE = sqrt(sq_sum / num_steps) (:449-450). */
int ddqc_ddmpc_action(const DdqcMpcConfig* cfg_host, const double* u_opt, int32_t n_step,
                      const float* bounds, float* action, double* u_applied, int64_t batch, void* stream);
int ddqc_ddmpc_post_step(const DdqcMpcConfig* cfg_host, double* u_win, double* y_win,
                         const double* u_applied, const float* y_meas, int64_t y_rs, int64_t y_cs,
                         const double* y_r, const double* d0, double max_inc, double* sq_sum,
                         int32_t* n_acc, int32_t* fail_step, double* y_log, int32_t step,
                         int64_t batch, void* gram_cache, void* stream);

/* Developer aid: device buffer of 12 int64 cycle counters accumulated per solver phase (NULL: off). */
int ddqc_ddmpc_debug_phase_buffer(void* dev_ptr);

/* ---------------------------------------------------------------------------------
 * Policy MLP forward (rsl_rl ActorCritic.act_inference; actor = Linear(16,128)-Tanh-Linear(128,128)-
 * Tanh-Linear(128,A), learning/config/hover_ppo_config.py:33-38, checkpoint layout of
 * models/demo_ctbr_fixed_yaw/model_1500.pt: actor.{0,2,4}.{weight,bias}).  Weights are PyTorch
 * Linear weights, [out][in] row-major fp32; obs [n][16] row-major fp32 (16-byte aligned); actions
 * [n][n_out] fp32.  precision 0: bf16 hi/lo split products (within 1e-4 of the fp32 forward); 1: single bf16
 * product.  Only n_in = 16, n_hidden = 128, n_out <= 4 is supported (DDQC_E_SHAPE otherwise).
 * --------------------------------------------------------------------------------- */
int ddqc_policy_mlp_forward(const float* obs, int64_t n, int32_t n_in, int32_t n_hidden, int32_t n_out,
                            const float* W1, const float* b1, const float* W2, const float* b2,
                            const float* W3, const float* b3, int32_t precision, float* actions,
                            void* stream);

/* library identification and layout self-check: sizeof of struct `which`
 * (0 DdqcEnvParams, 1 DdqcEnvCtl, 2 DdqcEnvState, 3 DdqcEnvOut, 4 DdqcMpcConfig, 5 DdqcMpcAux), -1 otherwise. */
int ddqc_abi_version(void);
int64_t ddqc_struct_size(int which);

#ifdef __cplusplus
}
#endif
#endif /* DDQC_H */
