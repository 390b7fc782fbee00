"""`Rotation3D`: typed wrapper around a rotation tensor, the interface every
model/loss of the reference exchanges (reference utils/rotation.py:91-309).

Same constructor, properties and tensor-like helpers; rotations are always
float32 (reference :141) and zero-norm (padded) quaternions become the identity
(reference :121-128)."""
import numpy as np
import torch

from . import rotation_conversions as _rc
from .rotation_conversions import rotation_6d_to_matrix as rot6d_to_matrix

EPS = 1e-6


def qeuler(q, order, epsilon=0, to_degree=False):
    """Quaternion(s) (*, 4) -> Euler angles (*, 3) (reference rotation.py:35-88)."""
    assert q.shape[-1] == 4
    out_shape = list(q.shape[:-1]) + [3]
    q0, q1, q2, q3 = q.reshape(-1, 4).unbind(1)
    lo, hi = -1 + epsilon, 1 - epsilon

    def asin(v):
        return torch.asin(torch.clamp(v, lo, hi))

    at2 = torch.atan2
    table = {
        'xyz': lambda: (at2(2 * (q0 * q1 - q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2)),
                        asin(2 * (q1 * q3 + q0 * q2)),
                        at2(2 * (q0 * q3 - q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))),
        'yzx': lambda: (at2(2 * (q0 * q1 - q2 * q3), 1 - 2 * (q1 * q1 + q3 * q3)),
                        at2(2 * (q0 * q2 - q1 * q3), 1 - 2 * (q2 * q2 + q3 * q3)),
                        asin(2 * (q1 * q2 + q0 * q3))),
        'zxy': lambda: (asin(2 * (q0 * q1 + q2 * q3)),
                        at2(2 * (q0 * q2 - q1 * q3), 1 - 2 * (q1 * q1 + q2 * q2)),
                        at2(2 * (q0 * q3 - q1 * q2), 1 - 2 * (q1 * q1 + q3 * q3))),
        'xzy': lambda: (at2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q3 * q3)),
                        at2(2 * (q0 * q2 + q1 * q3), 1 - 2 * (q2 * q2 + q3 * q3)),
                        asin(2 * (q0 * q3 - q1 * q2))),
        'yxz': lambda: (asin(2 * (q0 * q1 - q2 * q3)),
                        at2(2 * (q1 * q3 + q0 * q2), 1 - 2 * (q1 * q1 + q2 * q2)),
                        at2(2 * (q1 * q2 + q0 * q3), 1 - 2 * (q1 * q1 + q3 * q3))),
        'zyx': lambda: (at2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2)),
                        asin(2 * (q0 * q2 - q1 * q3)),
                        at2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))),
    }
    if order not in table:
        raise ValueError(f'unknown euler order {order}')
    euler = torch.stack(table[order](), dim=1).view(out_shape)
    if to_degree:
        euler = euler * 180. / np.pi
    return euler


_CONVERT = {
    ('quaternion', 'matrix'): _rc.quaternion_to_matrix,
    ('quaternion', 'axis_angle'): _rc.quaternion_to_axis_angle,
    ('matrix', 'quaternion'): _rc.matrix_to_quaternion,
    ('matrix', 'axis_angle'): _rc.matrix_to_axis_angle,
    ('axis_angle', 'quaternion'): _rc.axis_angle_to_quaternion,
    ('axis_angle', 'matrix'): _rc.axis_angle_to_matrix,
}


def _delegate(name, revalidate=True):
    """Tensor method applied to the wrapped rotation, re-wrapped.  Methods that
    keep values, dtype and the trailing representation axis (detach, clone,
    contiguous, cuda) skip the validity pass: the source was validated already
    (the reference re-runs it, ~5 small kernels per call, with the same result)."""

    def method(self, *args, **kwargs):
        out = getattr(self._rot, name)(*args, **kwargs)
        if revalidate:
            return type(self)(out, self._rot_type)
        new = object.__new__(type(self))
        new._rot, new._rot_type = out, self._rot_type
        return new

    method.__name__ = name
    return method


class Rotation3D:
    """3D rotation in one of three representations:
      'quat' (..., 4) real part first; 'rmat' (..., 3, 3) (6D input (..., 6) or
      (..., 2, 3) is converted in the constructor); 'axis' (..., 3)."""

    ROT_TYPE = ['quat', 'rmat', 'axis']
    ROT_NAME = {'quat': 'quaternion', 'rmat': 'matrix', 'axis': 'axis_angle'}

    def __init__(self, rot, rot_type='quat'):
        self._rot = rot
        self._rot_type = rot_type
        self._check_valid()

    def _process_zero_quat(self):
        r = self._rot
        if r.is_cuda and r.dtype == torch.float32 and not (torch.is_grad_enabled() and r.requires_grad):
            # one native launch (csrc/se3.cu) instead of norm / compare / zeros / fill / where
            from .. import _lib
            src = r.detach().contiguous()
            out = torch.empty_like(src)
            with torch.cuda.device(r.device):
                rc = _lib.lib().mpa_quat_fix_zero(_lib.ptr(src), src.numel() // 4, _lib.ptr(out),
                                                  _lib.cuda_stream(r.device))
            _lib.check(rc, 'mpa_quat_fix_zero')
            self._rot = out
            return
        with torch.no_grad():
            keep = torch.norm(self._rot, p=2, dim=-1, keepdim=True) > 0.5
            identity = torch.zeros_like(self._rot)
            identity[..., 0] = 1.
        self._rot = torch.where(keep, self._rot, identity)

    def _check_valid(self):
        assert self._rot_type in self.ROT_TYPE, \
            f'rotation {self._rot_type} is not supported'
        assert isinstance(self._rot, torch.Tensor), 'rotation must be a tensor'
        self._rot = self._rot.float()
        shape = self._rot.shape
        if self._rot_type == 'quat':
            assert shape[-1] == 4, 'wrong quaternion shape'
            self._process_zero_quat()
        elif self._rot_type == 'rmat':
            if shape[-1] == 3 and shape[-2] == 3:
                pass
            elif shape[-1] == 3 and shape[-2] == 2:
                self._rot = rot6d_to_matrix(self._rot.flatten(-2, -1))
            elif shape[-1] == 6:
                self._rot = rot6d_to_matrix(self._rot)
            elif shape[-1] == 3:
                raise ValueError('wrong rotation matrix shape')
            else:
                raise NotImplementedError('wrong rotation matrix shape')
        else:
            assert shape[-1] == 3

    def apply_rotation(self, rot):
        """Left-multiply by `rot`."""
        assert rot.rot_type in ['quat', 'rmat']
        rot = rot.convert(self._rot_type)
        if self._rot_type == 'quat':
            new_rot = _rc.quaternion_multiply(rot.rot, self._rot)
        else:
            new_rot = rot.rot @ self._rot
        return type(self)(new_rot, self._rot_type)

    def convert(self, rot_type):
        assert rot_type in self.ROT_TYPE, f'unknown target rotation {rot_type}'
        src, dst = self.ROT_NAME[self._rot_type], self.ROT_NAME[rot_type]
        if src == dst:
            return self.clone()
        return type(self)(_CONVERT[(src, dst)](self._rot), rot_type)

    def to_quat(self):
        return self.convert('quat').rot

    def to_rmat(self):
        return self.convert('rmat').rot

    def to_axis_angle(self):
        return self.convert('axis').rot

    def to_euler(self, order='zyx', to_degree=True):
        return qeuler(self.convert('quat')._rot, order=order, to_degree=to_degree)

    @property
    def rot(self):
        return self._rot

    @rot.setter
    def rot(self, rot):
        self._rot = rot
        self._check_valid()

    @property
    def rot_type(self):
        return self._rot_type

    @rot_type.setter
    def rot_type(self, rot_type):
        raise NotImplementedError(
            'please use convert() for rotation type conversion')

    @property
    def shape(self):
        return self._rot.shape

    @property
    def device(self):
        return self._rot.device

    @property
    def dtype(self):
        return self._rot.dtype

    reshape = _delegate('reshape')
    view = _delegate('view')
    flatten = _delegate('flatten')
    unflatten = _delegate('unflatten')
    transpose = _delegate('transpose')
    permute = _delegate('permute')
    contiguous = _delegate('contiguous', revalidate=False)
    to = _delegate('to')
    cuda = _delegate('cuda', revalidate=False)
    type = _delegate('type')
    type_as = _delegate('type_as')
    detach = _delegate('detach', revalidate=False)
    clone = _delegate('clone', revalidate=False)

    def squeeze(self, dim=None):
        r = self._rot.squeeze() if dim is None else self._rot.squeeze(dim)
        return self.__class__(r, self._rot_type)

    def unsqueeze(self, dim=None):
        return self.__class__(self._rot.unsqueeze(dim), self._rot_type)

    @staticmethod
    def _same_type(rot_lst):
        assert isinstance(rot_lst, (list, tuple))
        assert all(isinstance(r, Rotation3D) for r in rot_lst)
        rot_type = rot_lst[0].rot_type
        assert all(r.rot_type == rot_type for r in rot_lst)
        return rot_type, [r.rot for r in rot_lst]

    @staticmethod
    def cat(rot_lst, dim=0):
        rot_type, tensors = Rotation3D._same_type(rot_lst)
        return Rotation3D(torch.cat(tensors, dim=dim), rot_type)

    @staticmethod
    def stack(rot_lst, dim=0):
        rot_type, tensors = Rotation3D._same_type(rot_lst)
        return Rotation3D(torch.stack(tensors, dim=dim), rot_type)

    def __getitem__(self, key):
        return self.__class__(self._rot[key], self._rot_type)

    def __len__(self):
        return self._rot.shape[0]
