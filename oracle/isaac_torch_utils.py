"""TEST INFRASTRUCTURE ONLY -- restatement of the six ``isaacgym.torch_utils`` helpers the hot path uses.

Isaac Gym (NVIDIA, "Preview 3", unpinned; /root/reference/setup.py:11, README.md:17) is a closed
binary package that is NOT under /root/reference.  Its ``torch_utils.py`` ships as Python source
with the package; the formulas below restate that published file from the call sites the reference
makes (legged_gym/envs/base/legged_robot.py:37, 119-121, 338, 353-366, 405, 425, 431, 442,
536-538; legged_gym/utils/math.py:34).  PARITY UNPINNED for these six functions: the reference
holds no test or golden vector for them, so they are the de-facto spec.

The op sequence (mul / cross / bmm / norm / clamp) is kept identical to the published file so
that CPU rounding matches what the reference would produce with the real package installed.
"""
import torch

__all__ = ["quat_rotate_inverse", "quat_apply", "normalize", "torch_rand_float", "to_torch",
           "get_axis_params"]


def quat_rotate_inverse(q, v):
    # q = (x, y, z, w); rotate v by the inverse of q
    n = q.shape[0]
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(n, 1, 3), v.view(n, 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def quat_apply(a, b):
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = xyz.cross(b, dim=-1) * 2
    return (b + a[:, 3:] * t + xyz.cross(t, dim=-1)).view(shape)


def normalize(x, eps: float = 1e-9):
    return x / x.norm(p=2, dim=-1).clamp(min=eps, max=None).unsqueeze(-1)


def torch_rand_float(lower, upper, shape, device):
    return (upper - lower) * torch.rand(*shape, device=device) + lower


def to_torch(x, dtype=torch.float, device="cuda:0", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
    zs = [0.0] * n_dims
    zs[axis_idx] = 1.0
    params = [z * value for z in zs]
    params[0] = x_value if axis_idx != 0 else params[0]
    return [dtype(p) for p in params]
