"""torch helpers kept for user subclasses (mirror of reference legged_gym/utils/math.py:38-48 plus the
isaacgym.torch_utils functions user reward code calls).  The hot path itself does NOT use these: the
kernels in csrc/ implement the same formulas."""
import numpy as np
import torch


def quat_rotate_inverse(q, v):
    q_w = q[:, -1:]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0)
    b = torch.cross(q_vec, v, dim=-1) * q_w * 2.0
    c = q_vec * (q_vec * v).sum(-1, keepdim=True) * 2.0
    return a - b + c


def quat_apply(a, b):
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = torch.cross(xyz, b, dim=-1) * 2
    return (b + a[:, 3:] * t + torch.cross(xyz, t, dim=-1)).view(shape)


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)


def quat_apply_yaw(quat, vec):
    qy = quat.clone().view(-1, 4)
    qy[:, :2] = 0.
    return quat_apply(normalize(qy), vec)


def wrap_to_pi(angles):
    angles %= 2 * np.pi
    angles -= 2 * np.pi * (angles > np.pi)
    return angles
