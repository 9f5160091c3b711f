"""Restatement of the Isaac Gym Preview 4 `isaacgym.torch_utils` helpers used on the hot path.

TEST INFRASTRUCTURE (oracle).  Isaac Gym is a closed NVIDIA tarball (`setup.py:43` of the
reference pins it only as "isaacgym  # preview4"); it is not under /root/reference, so its
published algorithm is restated here.  PARITY UNPINNED for these helpers: the reference
holds no test for them.  Independent cross-check: reference `humanoid/scripts/sim2sim.py:57-77`
(`quaternion_to_euler_array`) and `:87-89` (scipy inverse rotation) -- see
tests/test_oracle_cpu.py::test_isaac_math_vs_sim2sim_formulas.

Call sites in the reference: `humanoid_env.py:51-56` (get_euler_xyz), `:275-277,785-787`
(quat_rotate_inverse).  Quaternions are (x, y, z, w).
"""
import math

import torch


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Rotate v [N,3] by the inverse of unit quaternion q [N,4] (xyzw).

    v*(2w^2-1) - 2w*(q_xyz x v) + 2*q_xyz*(q_xyz . v), in that op order.
    """
    w = q[:, 3]
    qv = q[:, :3]
    a = v * (2.0 * w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(qv, v, dim=-1) * w.unsqueeze(-1) * 2.0
    dot = torch.bmm(qv.reshape(-1, 1, 3), v.reshape(-1, 3, 1)).squeeze(-1)
    c = qv * dot * 2.0
    return a - b + c


def get_euler_xyz(q: torch.Tensor):
    """roll, pitch, yaw of q [N,4] (xyzw), each wrapped to [0, 2pi)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    sinr_cosp = 2.0 * (w * x + y * z)
    cosr_cosp = w * w - x * x - y * y + z * z
    roll = torch.atan2(sinr_cosp, cosr_cosp)
    sinp = 2.0 * (w * y - z * x)
    half_pi = torch.full_like(sinp, math.pi / 2.0)
    pitch = torch.where(torch.abs(sinp) >= 1, torch.copysign(half_pi, sinp), torch.asin(sinp))
    siny_cosp = 2.0 * (w * z + x * y)
    cosy_cosp = w * w + x * x - y * y - z * z
    yaw = torch.atan2(siny_cosp, cosy_cosp)
    two_pi = 2 * math.pi
    return roll % two_pi, pitch % two_pi, yaw % two_pi


def euler_xyz_wrapped(q: torch.Tensor) -> torch.Tensor:
    """`get_euler_xyz_tensor` of the reference (`humanoid_env.py:51-56`): stack, then
    subtract 2pi where > pi so every angle lies in (-pi, pi]."""
    r, p, y = get_euler_xyz(q)
    e = torch.stack((r, p, y), dim=1)
    e[e > math.pi] -= 2 * math.pi
    return e


def quat_apply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Rotate b [N,3] by unit quaternion a [N,4] (xyzw): b + w*t + q_xyz x t with t = 2*(q_xyz x b), in that op order
    (Isaac Gym Preview 4 `torch_utils.quat_apply`; call site `humanoid_env.py:1008`, heading of the base)."""
    shape = b.shape
    a = a.reshape(-1, 4)
    b = b.reshape(-1, 3)
    xyz = a[:, :3]
    t = torch.cross(xyz, b, dim=-1) * 2
    return (b + a[:, 3:] * t + torch.cross(xyz, t, dim=-1)).view(shape)


def to_torch(x, dtype=torch.float, device="cpu", requires_grad=False):
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
    zs = [0.0] * n_dims
    zs[axis_idx] = 1.0
    params = [z * value for z in zs]
    params[0] = x_value if axis_idx != 0 else params[0]
    return params
