"""SE(3) transforms of point clouds, reference API of utils/transforms.py.

`qrot` / `qtransform` (reference :75-109) are the hot functions: the reference
materialises the per-part quaternion to [B,P,N,4] with repeat_interleave and
runs pytorch3d's quaternion_apply (~40 elementwise kernels); here they are one
CUDA kernel (csrc/se3.cu) that keeps q,t in registers per part, with a matching
backward kernel.  Rotation-matrix variants stay small broadcasted matmuls.
"""
import numpy as np
import torch

from .. import _lib
from .rotation import Rotation3D
from .rotation_conversions import quaternion_invert, quaternion_raw_multiply, \
    quaternion_to_matrix, matrix_to_quaternion
from .rotation_conversions import random_quaternions as _random_quaternions


def random_quaternions(shape):
    """Random unit quaternions with non-negative real part, shape [..., 4]."""
    assert isinstance(shape, (int, list, tuple))
    shape = [shape] if isinstance(shape, int) else list(shape)
    quat = _random_quaternions(int(np.prod(shape)))
    return quat.view(shape + [4])


def qmul(q, r):
    """Hamilton product of (*, 4) quaternions."""
    return quaternion_raw_multiply(q, r)


def qrmat(q):
    """Quaternion(s) (*, 4) -> rotation matrix (*, 3, 3)."""
    assert q.shape[-1] == 4
    return quaternion_to_matrix(q)


class _SE3Function(torch.autograd.Function):
    """out[p, i] = q_p (x) (0, pts[p, i]) (x) conj(q_p) (+ t_p); fp32 CUDA."""

    @staticmethod
    def forward(ctx, quat, trans, pts):
        # quat [n,4], trans [n,3] or None, pts [n,N,3]; all contiguous fp32 CUDA
        n, N, _ = pts.shape
        out = torch.empty_like(pts)
        with torch.cuda.device(pts.device):
            rc = _lib.lib().mpa_se3_transform(
                _lib.ptr(quat), _lib.ptr(trans), _lib.ptr(pts), n, N,
                _lib.ptr(out), _lib.cuda_stream(pts.device))
        _lib.check(rc, 'mpa_se3_transform')
        ctx.save_for_backward(quat, pts)
        ctx.has_trans = trans is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        quat, pts = ctx.saved_tensors
        n, N, _ = pts.shape
        grad_out = grad_out.contiguous().float()
        need_q, need_t, need_p = ctx.needs_input_grad
        need_t = need_t and ctx.has_trans
        gq = torch.empty_like(quat) if need_q else None
        gt = torch.empty(n, 3, dtype=torch.float32, device=pts.device) if need_t else None
        gp = torch.empty_like(pts) if need_p else None
        with torch.cuda.device(pts.device):
            rc = _lib.lib().mpa_se3_transform_backward(
                _lib.ptr(quat), _lib.ptr(pts), _lib.ptr(grad_out), n, N,
                _lib.ptr(gp), _lib.ptr(gq), _lib.ptr(gt),
                _lib.cuda_stream(pts.device))
        _lib.check(rc, 'mpa_se3_transform_backward')
        return gq, gt, gp


def _se3(q, t, v):
    """Shared driver of qrot/qtransform with the reference's broadcasting rule:
    q (and t) either match v's leading shape, or lack the point axis (-2)."""
    _lib.require_cuda(q, v, t)
    assert q.shape[-1] == 4 and v.shape[-1] == 3
    if q.dim() == v.dim() - 1:
        assert q.shape[:-1] == v.shape[:-2]
        N = v.shape[-2]
    else:
        assert q.shape[:-1] == v.shape[:-1]
        N = 1
    if t is not None:
        assert t.shape[-1] == 3 and t.shape[:-1] == q.shape[:-1]
        t = t.reshape(-1, 3).contiguous().float()
    out = _SE3Function.apply(
        q.reshape(-1, 4).contiguous().float(), t,
        v.reshape(-1, N, 3).contiguous().float())
    return out.view(v.shape)


def qrot(q, v):
    """Rotate v (*, 3) by q (*, 4); q may omit the point axis, e.g.
    [B, P, 4] with v [B, P, N, 3] (reference :75-87)."""
    return _se3(q, None, v)


def qtransform(t, q, v):
    """Rotate v by q then translate by t (reference :90-109)."""
    assert t.shape[-1] == 3
    if t.dim() == v.dim() - 1 and q.dim() == v.dim():
        # per-point quaternion with a per-cloud translation: expand t
        t = t.unsqueeze(-2).expand(v.shape)
    elif t.dim() == v.dim() and q.dim() == v.dim() - 1:
        q = q.unsqueeze(-2).expand(v.shape[:-1] + (4, ))
    return _se3(q, t, v)


def qtransform_invert(t, q, tqv):
    """Inverse of qtransform (reference :112-124)."""
    assert t.shape[-1] == 3
    if t.dim() == tqv.dim() - 1:
        t = t.unsqueeze(-2)
    return qrot(quaternion_invert(q), tqv - t)


def random_rotation_matrixs(shape):
    return quaternion_to_matrix(random_quaternions(shape))


def rmatq(r):
    assert r.shape[-1] == r.shape[-2] == 3
    return matrix_to_quaternion(r)


def rmat_rot(r, v):
    """Rotate v (*, 3) by rotation matrices r (*, 3, 3) (reference :155-172);
    broadcast instead of the reference's repeat_interleave."""
    assert r.shape[-1] == r.shape[-2] == 3 and v.shape[-1] == 3
    if r.dim() == v.dim():
        r = r.unsqueeze(-3)
    assert r.shape[:-3] == v.shape[:-2]
    return (r @ v.unsqueeze(-1)).squeeze(-1)


def rmat_transform(t, r, v):
    """Rotate by r then translate by t (reference :175-194)."""
    assert t.shape[-1] == 3
    if t.dim() == v.dim() - 1:
        t = t.unsqueeze(-2)
    return rmat_rot(r, v) + t


def _unwrap(rot, rot_type):
    if rot_type is None:
        assert isinstance(rot, Rotation3D)
        return rot.rot, rot.rot_type
    assert isinstance(rot, torch.Tensor)
    return rot, rot_type


def rot_pc(rot, pc, rot_type=None):
    """Rotate a point cloud by a Rotation3D (or a tensor + rot_type)
    (reference :199-220)."""
    r, rot_type = _unwrap(rot, rot_type)
    if rot_type == 'quat':
        return qrot(r, pc)
    elif rot_type == 'rmat':
        return rmat_rot(r, pc)
    raise NotImplementedError(f'{rot_type} is not supported')


def transform_pc(trans, rot, pc, rot_type=None):
    """Rotate then translate a point cloud (reference :223-244)."""
    r, rot_type = _unwrap(rot, rot_type)
    if rot_type == 'quat':
        return qtransform(trans, r, pc)
    elif rot_type == 'rmat':
        return rmat_transform(trans, r, pc)
    raise NotImplementedError(f'{rot_type} is not supported')


def quaternion_to_rmat(quat):
    """quat [4] (w, i, j, k) numpy -> 3x3 numpy."""
    from scipy.spatial.transform import Rotation as R
    return R.from_quat(quat[[1, 2, 3, 0]]).as_matrix()


def trans_rmat_to_pmat(trans, rmat):
    pose_mat = np.eye(4)
    pose_mat[:3, :3] = rmat
    pose_mat[:3, -1] = trans
    return pose_mat


def trans_quat_to_pmat(trans, quat):
    return trans_rmat_to_pmat(trans, quaternion_to_rmat(quat))
