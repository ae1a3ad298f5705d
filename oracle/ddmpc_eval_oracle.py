"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy restatement of the reference's closed-loop evaluation of
one data-driven MPC controller, `sim_nonlinear_dd_mpc_control_loop_parallel`
(data_driven_mpc/utilities/param_grid_search/controller_evaluation.py:345-450), with the
multiprocessing queues replaced by a `step_env(u) -> y` callable.  Pinned by the reference text
itself: same loop structure, RMSE formula (:449-450) and early-stop rule (:424-447)."""

from __future__ import annotations

import math

import numpy as np


class MovedAway(ValueError):
    pass


def sim_control_loop(controller, step_env, num_steps, initial_distance, max_target_dist_increment=None, n_n_mpc_step=False):
    n, y_r = controller.n, np.asarray(controller.y_r, float).reshape(-1, 1)
    n_mpc_step = n if n_n_mpc_step else 1  # :351
    u0 = np.asarray(controller.get_optimal_control_input_at_step(0) if controller.optimal_u is not None else np.zeros(controller.m))
    m, p = u0.shape[0], y_r.shape[0]
    u_sys, y_sys = np.zeros((num_steps, m)), np.zeros((num_steps, p))
    for t in range(0, num_steps, n_mpc_step):  # :356
        controller.update_and_solve_data_driven_mpc()  # :361
        for k in range(t, min(t + n_mpc_step, num_steps)):  # :364
            n_step = k - t
            u_sys[k] = controller.get_optimal_control_input_at_step(n_step=n_step)  # :389-397
            y_sys[k] = step_env(u_sys[k])  # :400-409
            controller.store_input_output_measurement(u_current=u_sys[k], y_current=y_sys[k], du_current=None)  # :416-422
            if max_target_dist_increment is not None:  # :424-447
                current_distance = np.linalg.norm(y_sys[k].reshape(-1, 1) - y_r)
                if current_distance - initial_distance > max_target_dist_increment:
                    raise MovedAway(k)
    return math.sqrt(np.mean(np.sum((y_sys - y_r.T) ** 2, axis=1))), u_sys, y_sys  # :449-450
