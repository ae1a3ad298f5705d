"""Env construction and state helpers for callers of the hot path
(reference: utilities/drone_environment.py:12-80)."""

from typing import Any

import torch

from data_driven_quad_control_b200.envs.hover_env import HoverEnv
from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType, EnvState

EnvIdx = int | list[int] | tuple[int, ...] | torch.Tensor


def create_env(
    num_envs: int,
    env_cfg: Any,
    obs_cfg: Any,
    reward_cfg: Any,
    command_cfg: Any,
    show_viewer: bool = False,
    device: torch.device | str = "cuda",
    action_type: EnvActionType = EnvActionType.CTBR_FIXED_YAW,
    drone_colors: list[tuple[float, float, float, float]] | None = None,
    camera_config: dict[str, Any] | None = None,
    **kwargs: Any,
) -> HoverEnv:
    """HoverEnv with automatic target updates disabled (targets are set by the caller)."""
    return HoverEnv(
        num_envs=num_envs, env_cfg=env_cfg, obs_cfg=obs_cfg, reward_cfg=reward_cfg, command_cfg=command_cfg,
        show_viewer=show_viewer, device=device, action_type=action_type, drone_colors=drone_colors,
        camera_config=camera_config, auto_target_updates=False, **kwargs,
    )  # fmt: skip


def get_tensor_from_env_idx(env: HoverEnv, env_idx: EnvIdx) -> torch.Tensor:
    if isinstance(env_idx, int):
        return torch.tensor([env_idx], dtype=torch.long, device=env.device)
    if isinstance(env_idx, (list, tuple)):
        return torch.tensor(env_idx, dtype=torch.long, device=env.device)
    if isinstance(env_idx, torch.Tensor):
        return env_idx.to(dtype=torch.long, device=env.device)
    raise TypeError(f"Unsupported env_idx type: {type(env_idx)}")


def get_current_env_state(env: HoverEnv, env_idx: int) -> EnvState:
    return env.get_current_state(envs_idx=get_tensor_from_env_idx(env, env_idx))


def restore_env_from_state(env: HoverEnv, env_idx: int, saved_state: EnvState) -> None:
    env.restore_from_state(envs_idx=get_tensor_from_env_idx(env, env_idx), saved_state=saved_state)


def update_env_target_pos(env: HoverEnv, env_idx: EnvIdx, target_pos: torch.Tensor) -> None:
    env.update_target_pos(envs_idx=get_tensor_from_env_idx(env, env_idx), target_pos=target_pos)
