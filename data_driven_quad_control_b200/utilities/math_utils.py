"""Small tensor helpers used by the callers of the hot path (reference: utilities/math_utils.py).

These run on whatever device their inputs live on; they are glue for callers (normalising a
controller command into an env action, building a yaw quaternion), not part of the fused path.
"""

import torch


def gs_rand_float(lower: float, upper: float, shape: tuple[int, ...], device: torch.device) -> torch.Tensor:
    """Uniform samples in [lower, upper)  (math_utils.py:4-10)."""
    return (upper - lower) * torch.rand(*shape, device=device) + lower


def linear_interpolate(x, x_min, x_max, y_min, y_max):
    """Affine map of [x_min, x_max] onto [y_min, y_max], op order of math_utils.py:13-20."""
    return y_min + (x - x_min) * (y_max - y_min) / (x_max - x_min)


def quaternion_to_matrix(quats: torch.Tensor) -> torch.Tensor:
    """(…,4) w-x-y-z quaternions -> (…,3,3) rotation matrices (math_utils.py:23-61)."""
    w, x, y, z = quats.unbind(-1)
    ww, xx, yy, zz = w * w, x * x, y * y, z * z
    xy, xz, yz, wx, wy, wz = x * y, x * z, y * z, w * x, w * y, w * z
    rows = (
        ww + xx - yy - zz, 2 * (xy - wz), 2 * (xz + wy),
        2 * (xy + wz), ww - xx + yy - zz, 2 * (yz - wx),
        2 * (xz - wy), 2 * (yz + wx), ww - xx - yy + zz,
    )  # fmt: skip
    return torch.stack(rows, dim=-1).reshape(quats.shape[:-1] + (3, 3))


def yaw_from_quaternion(quats: torch.Tensor) -> torch.Tensor:
    """ZYX Tait-Bryan yaw of (…,4) quaternions, radians (math_utils.py:64-83)."""
    w, x, y, z = quats.unbind(-1)
    return torch.atan2(2 * (w * z + x * y), 1 - 2 * (y**2 + z**2))


def yaw_to_quaternion(yaw: torch.Tensor) -> torch.Tensor:
    """Pure-yaw quaternions (w, 0, 0, z) for a 1-D tensor of yaw angles (math_utils.py:86-107)."""
    quat = torch.zeros((yaw.shape[0], 4), device=yaw.device, dtype=yaw.dtype)
    quat[:, 0] = torch.cos(yaw / 2)
    quat[:, 3] = torch.sin(yaw / 2)
    return quat
