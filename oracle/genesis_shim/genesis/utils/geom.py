"""TEST INFRASTRUCTURE ONLY -- the four `genesis.utils.geom` helpers the reference imports
(`hover_env.py:44-49`), restated for torch tensors (w,x,y,z quaternions)."""

import math

import torch

import drone_physics as _phys


def inv_quat(quat):
    out = quat.clone()
    out[..., 1:] = -out[..., 1:]
    return out


def transform_by_quat(v, quat):
    x, y, z = _phys.rotate_by_quat(*v.unbind(-1), *quat.unbind(-1))
    return torch.stack((x, y, z), -1)


def transform_quat_by_quat(v, u):
    """Apply v first, then u:  u (x) v."""
    w, x, y, z = _phys.quat_mul(*u.unbind(-1), *v.unbind(-1))
    return torch.stack((w, x, y, z), -1)


def quat_to_xyz(quat, rpy=False, degrees=False):
    qw, qx, qy, qz = quat.unbind(-1)
    roll = torch.atan2((qw * qx + qy * qz) * 2.0, 1.0 - (qx * qx + qy * qy) * 2.0)
    pitch = torch.asin(torch.clamp((qw * qy - qz * qx) * 2.0, -1.0, 1.0))
    yaw = torch.atan2((qw * qz + qx * qy) * 2.0, 1.0 - (qy * qy + qz * qz) * 2.0)
    out = torch.stack((roll, pitch, yaw), -1)
    if degrees:
        out = out * (180.0 / math.pi)
    return out
