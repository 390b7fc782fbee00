"""Rotation-representation conversions with the semantics of the pytorch3d
functions the reference imports (utils/transforms.py:19-24, utils/rotation.py:6-10).
pytorch3d is an un-vendored, unpinned third-party dependency of the reference;
these are restated from its published behaviour (real-part-first quaternions).

Only `quaternion_apply` is on the hot path, and that one is the CUDA kernel in
csrc/se3.cu (see transforms.qrot); everything here is small [B,P,*] glue.
"""
import torch
import torch.nn.functional as F


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ), -1)


def standardize_quaternion(q):
    return torch.where(q[..., :1] < 0, -q, q)


def quaternion_multiply(a, b):
    return standardize_quaternion(quaternion_raw_multiply(a, b))


def quaternion_invert(q):
    return q * q.new_tensor([1.0, -1.0, -1.0, -1.0])


def random_quaternions(n, dtype=None, device=None):
    o = torch.randn((n, 4), dtype=dtype, device=device)
    s = (o * o).sum(1)
    return o / torch.copysign(torch.sqrt(s), o[:, 0])[:, None]


def quaternion_to_matrix(q):
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    m = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
    ), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(matrix):
    if matrix.shape[-1] != 3 or matrix.shape[-2] != 3:
        raise ValueError(f'Invalid rotation matrix shape {matrix.shape}.')
    batch = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = \
        matrix.reshape(batch + (9, )).unbind(-1)
    q_abs = _sqrt_positive_part(torch.stack([
        1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
        1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    cand = torch.stack([
        torch.stack([q_abs[..., 0]**2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1]**2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2]**2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3]**2], -1),
    ], dim=-2)
    floor = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * q_abs[..., None].max(floor))
    best = F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return standardize_quaternion(cand[best, :].reshape(batch + (4, )))


def axis_angle_to_quaternion(axis_angle):
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = angles * 0.5
    small = angles.abs() < 1e-6
    k = torch.empty_like(angles)
    k[~small] = torch.sin(half[~small]) / angles[~small]
    k[small] = 0.5 - (angles[small] * angles[small]) / 48
    return torch.cat([torch.cos(half), axis_angle * k], dim=-1)


def quaternion_to_axis_angle(q):
    norms = torch.norm(q[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    angles = 2 * half
    small = angles.abs() < 1e-6
    k = torch.empty_like(angles)
    k[~small] = torch.sin(half[~small]) / angles[~small]
    k[small] = 0.5 - (angles[small] * angles[small]) / 48
    return q[..., 1:] / k


def axis_angle_to_matrix(axis_angle):
    return quaternion_to_matrix(axis_angle_to_quaternion(axis_angle))


def matrix_to_axis_angle(matrix):
    return quaternion_to_axis_angle(matrix_to_quaternion(matrix))


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)
