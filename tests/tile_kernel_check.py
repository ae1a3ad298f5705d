"""Helper of tests/test_env_tile_kernel_gpu.py: runs in a subprocess whose environment forces the bulk-copy tile kernel
(DDQC_ENV_TILE_MIN_N small, DDQC_ENV_FUSED_MAX_N = 0) at sizes the numpy oracle steps in seconds."""

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE]

from test_env_parity_gpu import _actions, _compare_step, _make_pair  # noqa: E402


def run(at_name: str, N: int, T: int, auto: bool, independent: bool, park: bool) -> dict:
    env, orc = _make_pair(N, at_name, auto, materialize=False, independent=independent)
    acts = _actions(T, N, env.num_actions, 21)
    if park:
        acts[:, : N // 8] = 0.0  # hovering drones: hover counter, target resample, A15/A16 fix-ups
    env.reset()
    orc.reset()
    n_reset = n_hover = 0
    for t in range(T):
        if park and t % 20 == 0:
            idx = np.arange(N // 8)
            tgt = orc.base_pos[idx].copy()
            orc.update_target_pos(idx, tgt)
            env.update_target_pos(torch.arange(N // 8, device="cuda"), torch.from_numpy(tgt).cuda())
        if t == T // 2:  # an external reset between two tile-kernel steps
            idx = [1, N // 2, N - 1]
            env.reset_idx(torch.tensor(idx, device="cuda"))
            orc.reset_idx(idx)
        env.step(torch.from_numpy(acts[t]).cuda())
        orc.step(acts[t])
        _compare_step(env, orc, t, yaw_tol=0.0 if at_name == "CTBR_FIXED_YAW" else 2e-6)
        n_reset += int(orc.reset_buf.sum())
        n_hover += int(orc.hover_resampled.sum()) if hasattr(orc, "hover_resampled") else 0
        for k, v in orc.extras_episode.items():
            assert abs(float(env.extras["episode"][k]) - v) <= 1e-5 * max(1.0, abs(v)), (t, k)
    stats = env.rollout_stats()
    assert int(stats["resets"]) == n_reset and int(stats["env_steps"]) == N * T
    return {"resets": n_reset, "hover_events": n_hover}


if __name__ == "__main__":
    at_name, N, T, auto, independent, park = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), *(a == "1" for a in sys.argv[4:7])
    print(json.dumps(run(at_name, N, T, auto, independent, park)))
