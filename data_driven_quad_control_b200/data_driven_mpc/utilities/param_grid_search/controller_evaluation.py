"""Batched closed-loop evaluation of nonlinear data-driven MPC controllers on the GPU env.

Device-resident counterpart of the reference's `sim_nonlinear_dd_mpc_control_loop_parallel`
(data_driven_mpc/utilities/param_grid_search/controller_evaluation.py:283-450).  There, one CPU
process per controller exchanges actions / observations with the process that steps the
vectorized env through multiprocessing queues.  Here controller `b` drives env `b` of the same
`HoverEnv`, and one control step is four asynchronous launches on one stream with no host
synchronisation:

    ddqc_ddmpc_solve        update_and_solve_data_driven_mpc            (:361, every n_mpc_step steps)
    ddqc_ddmpc_action       get_optimal_control_input_at_step + action normalisation   (:389-397)
    ddqc_env_step           the main process' env.step                  (parallel_grid_search.py)
    ddqc_ddmpc_post_step    store_input_output_measurement (:416-422) + squared tracking error +
                            early stop |y - y_r| - d0 > max_target_dist_increment      (:424-447)

The RMSE is `sqrt(mean_t sum_axis (y - y_r)^2)` over the `num_steps` steps (:449-450).  A controller
whose drone moved away from its target by more than the allowed increment stops at that step
(the reference raises ValueError there, which its caller turns into a FAILURE evaluation,
:178-215): its windows are frozen, `failed[b]` is set and its RMSE is NaN.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from data_driven_quad_control_b200 import _lib
from data_driven_quad_control_b200.data_driven_mpc.nonlinear_data_driven_mpc_controller import (
    BatchedNonlinearDataDrivenMPC,
)
from data_driven_quad_control_b200.envs.hover_env import HoverEnv


@dataclass
class BatchedEvalResult:
    rmse: torch.Tensor  # (B,) float64, NaN where the evaluation stopped early
    failed: torch.Tensor  # (B,) bool: early stop (drone moved away) or solver failure
    fail_step: torch.Tensor  # (B,) int32, -1 where the run completed
    solver_failed: torch.Tensor  # (B,) bool: a solve returned a non-zero status at some step
    y_sys: torch.Tensor | None  # (num_steps, B, p) float64 when `record` is set
    u_sys: torch.Tensor | None  # (num_steps, B, m) float64 when `record` is set


def sim_nonlinear_dd_mpc_control_loop_batched(
    env: HoverEnv,
    dd_mpc_controller: BatchedNonlinearDataDrivenMPC,
    num_steps: int,
    initial_distance: torch.Tensor | None = None,
    max_target_dist_increment: float | None = None,
    n_n_mpc_step: bool = False,
    record: bool = False,
) -> BatchedEvalResult:
    """Closed loop of every controller of the batch on its own env for `num_steps` steps."""
    ctrl = dd_mpc_controller
    B, m, p = ctrl.B, ctrl.m, ctrl.p
    if env.num_envs != B:
        raise ValueError(f"controller batch ({B}) and env.num_envs ({env.num_envs}) must match")
    if env.num_actions != m:
        raise ValueError("the env action dimension must equal the controller input dimension m")
    n_mpc_step = ctrl.n if n_n_mpc_step else 1
    if n_mpc_step > ctrl.n_mpc_step:
        raise ValueError("the controller was built with n_mpc_step smaller than the requested step count per solve")
    dev = ctrl.device
    lib = _lib.load()
    f64 = dict(dtype=torch.float64, device=dev)
    y_r = ctrl.y_r
    if initial_distance is None:
        initial_distance = torch.linalg.norm(env.get_pos(add_noise=False).to(torch.float64) - y_r, dim=1)
    d0 = torch.as_tensor(initial_distance).to(**f64).reshape(B).contiguous()
    max_inc = -1.0 if max_target_dist_increment is None else float(max_target_dist_increment)
    sq_sum = torch.zeros((B,), **f64)
    n_acc = torch.zeros((B,), dtype=torch.int32, device=dev)
    fail_step = torch.full((B,), -1, dtype=torch.int32, device=dev)
    solver_failed = torch.zeros((B,), dtype=torch.bool, device=dev)
    action = torch.empty((B, m), dtype=torch.float32, device=dev)
    u_applied = torch.empty((B, m), **f64)
    bounds = env.action_bounds.to(torch.float32).contiguous()
    y_log = torch.zeros((num_steps, B, p), **f64) if record else None
    u_log = torch.zeros((num_steps, B, m), **f64) if record else None
    cfg = C.byref(ctrl._cfg)
    with _lib.device_guard(dev):  # the raw C calls of the loop launch on the current device
        return _closed_loop(env, ctrl, lib, cfg, num_steps, n_mpc_step, B, m, bounds, action, u_applied, y_r, d0, max_inc,
                            sq_sum, n_acc, fail_step, solver_failed, y_log, u_log, record, dev)  # fmt: skip


def _closed_loop(env, ctrl, lib, cfg, num_steps, n_mpc_step, B, m, bounds, action, u_applied, y_r, d0, max_inc, sq_sum, n_acc,
                 fail_step, solver_failed, y_log, u_log, record, dev):  # fmt: skip
    for t in range(0, num_steps, n_mpc_step):
        # runs that already stopped (drift criterion or a failed solve) are aborted evaluations in the reference
        # (controller_evaluation.py:424-447): their controllers are left out of the solve
        ctrl.update_and_solve_data_driven_mpc(skip=(fail_step >= 0) | solver_failed)
        solver_failed |= ctrl.status != 0
        for k in range(t, min(t + n_mpc_step, num_steps)):
            rc = lib.ddqc_ddmpc_action(cfg, ctrl.optimal_u.data_ptr(), k - t, bounds.data_ptr(), action.data_ptr(),
                                       u_applied.data_ptr(), B, _lib.stream_ptr(dev))  # fmt: skip
            _lib.check(rc, "ddqc_ddmpc_action")
            env.step(action)
            y = env.get_pos()  # the measured output: position (+ observation noise, as the reference's env)
            rc = lib.ddqc_ddmpc_post_step(
                cfg, ctrl.u.data_ptr(), ctrl.y.data_ptr(), u_applied.data_ptr(), y.data_ptr(), y.stride(0), y.stride(1),
                y_r.data_ptr(), d0.data_ptr(), max_inc, sq_sum.data_ptr(), n_acc.data_ptr(), fail_step.data_ptr(),
                y_log.data_ptr() if record else None, k, B, ctrl.gram_cache_ptr, _lib.stream_ptr(dev),
            )  # fmt: skip
            _lib.check(rc, "ddqc_ddmpc_post_step")
            if record:
                u_log[k] = u_applied
    failed = (fail_step >= 0) | solver_failed
    rmse = torch.sqrt(sq_sum / num_steps)
    rmse = torch.where(failed, torch.full_like(rmse, float("nan")), rmse)
    return BatchedEvalResult(rmse=rmse, failed=failed, fail_step=fail_step, solver_failed=solver_failed, y_sys=y_log, u_sys=u_log)
