"""Reward scales per env action type (reference: learning/config/reward_config.py:6-51)."""

from typing import Any

from data_driven_quad_control_b200.envs.hover_env_config import EnvActionType

_COMMON = {"target": 10.0, "closeness": 1.5, "hover_time": 0.01, "yaw": 0.01, "angular": -2e-4, "crash": -10.0}
_SMOOTH = {EnvActionType.ROTOR_RPMS: -1e-4, EnvActionType.CTBR: -0.01, EnvActionType.CTBR_FIXED_YAW: -0.01}


def get_reward_cfg_by_action_type(action_type: EnvActionType) -> dict[str, Any]:
    if action_type not in _SMOOTH:
        raise ValueError(f"Unsupported action type: {action_type}")
    scales = {
        "target": _COMMON["target"],
        "closeness": _COMMON["closeness"],
        "hover_time": _COMMON["hover_time"],
        "smooth": _SMOOTH[action_type],
        # yaw cannot be controlled when the yaw rate action is fixed
        "yaw": 0.0 if action_type == EnvActionType.CTBR_FIXED_YAW else _COMMON["yaw"],
        "angular": _COMMON["angular"],
        "crash": _COMMON["crash"],
    }
    return {"yaw_lambda": -10.0, "reward_scales": scales}
