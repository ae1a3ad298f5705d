"""TEST INFRASTRUCTURE ONLY -- the four `genesis.utils.geom` helpers the reference imports
(`hover_env.py:44-49`), restated for torch tensors (w,x,y,z quaternions)."""

import math

import numpy as np
import torch

import drone_physics as _phys


def inv_quat(quat):
    out = quat.clone()
    out[..., 1:] = -out[..., 1:]
    return out


def _cols(t):
    a = t.detach().cpu().numpy().astype(np.float32)
    return tuple(np.ascontiguousarray(a[..., j]) for j in range(a.shape[-1]))


def _stack(cols):
    return torch.from_numpy(np.stack([np.asarray(c, dtype=np.float32) for c in cols], -1))


def transform_by_quat(v, quat):
    return _stack(_phys.rotate_by_quat(*_cols(v), *_cols(quat)))


def transform_quat_by_quat(v, u):
    """Apply v first, then u:  u (x) v."""
    v, u = torch.broadcast_tensors(v, u)
    return _stack(_phys.quat_mul(*_cols(u), *_cols(v)))


def quat_to_xyz(quat, rpy=False, degrees=False):
    qw, qx, qy, qz = quat.unbind(-1)
    roll = torch.atan2((qw * qx + qy * qz) * 2.0, 1.0 - (qx * qx + qy * qy) * 2.0)
    pitch = torch.asin(torch.clamp((qw * qy - qz * qx) * 2.0, -1.0, 1.0))
    yaw = torch.atan2((qw * qz + qx * qy) * 2.0, 1.0 - (qy * qy + qz * qz) * 2.0)
    out = torch.stack((roll, pitch, yaw), -1)
    if degrees:
        out = out * (180.0 / math.pi)
    return out
