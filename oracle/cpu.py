"""ctypes/numpy front-end of oracle/libmpa_oracle.so (TEST INFRASTRUCTURE ONLY).

Each wrapper mirrors one native entry point of the reference:
  chamfer_forward / chamfer_backward -> utils/chamfer/cuda/chamfer.cpp:8-23
  se3_transform                      -> utils/transforms.py:75-109 (+pytorch3d)
  knn                                -> models/modules/encoder/dgcnn.py:8-15
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile oracle/libmpa_oracle.so with gcc (seconds)."""
    so = os.path.join(_HERE, 'libmpa_oracle.so')
    src = os.path.join(_HERE, 'mpa_oracle.c')
    if force or not os.path.exists(so) or \
            os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B', 'all'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def chamfer_forward(xyz1, xyz2, fused=True):
    """(B,N1,3),(B,N2,3) float32 -> dist1 f32 (B,N1), idx1 i64, dist2, idx2.

    fused=True follows the reference CUDA kernel bit for bit (nvcc FMA
    contraction of chamfer_kernel.cu:80); fused=False is the un-contracted
    brute-force definition of test_chamfer.py:8-31.
    """
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, n1, _ = xyz1.shape
    n2 = xyz2.shape[1]
    assert xyz2.shape[0] == B and xyz1.shape[2] == 3 and xyz2.shape[2] == 3
    d1 = np.empty((B, n1), np.float32)
    i1 = np.empty((B, n1), np.int64)
    d2 = np.empty((B, n2), np.float32)
    i2 = np.empty((B, n2), np.int64)
    lib().oracle_chamfer_forward(
        _p(xyz1, _f32p), _p(xyz2, _f32p), ctypes.c_int64(B),
        ctypes.c_int64(n1), ctypes.c_int64(n2), _p(d1, _f32p), _p(i1, _i64p),
        _p(d2, _f32p), _p(i2, _i64p), ctypes.c_int(1 if fused else 0))
    return d1, i1, d2, i2


def chamfer_backward(grad_dist1, grad_dist2, xyz1, xyz2, idx1, idx2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    g1, g2 = _f32(grad_dist1), _f32(grad_dist2)
    idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
    idx2 = np.ascontiguousarray(idx2, dtype=np.int64)
    B, n1, _ = xyz1.shape
    n2 = xyz2.shape[1]
    gx1 = np.empty_like(xyz1)
    gx2 = np.empty_like(xyz2)
    lib().oracle_chamfer_backward(
        _p(g1, _f32p), _p(g2, _f32p), _p(xyz1, _f32p), _p(xyz2, _f32p),
        _p(idx1, _i64p), _p(idx2, _i64p), ctypes.c_int64(B),
        ctypes.c_int64(n1), ctypes.c_int64(n2), _p(gx1, _f32p), _p(gx2, _f32p))
    return gx1, gx2


def se3_transform(quat, trans, pts):
    """quat [..., 4], trans [..., 3] or None, pts [..., N, 3] -> [..., N, 3]."""
    pts = _f32(pts)
    quat = _f32(quat)
    lead = pts.shape[:-2]
    N = pts.shape[-2]
    n_parts = int(np.prod(lead)) if lead else 1
    assert quat.shape == lead + (4, )
    out = np.empty_like(pts)
    tp = None
    if trans is not None:
        trans = _f32(trans)
        assert trans.shape == lead + (3, )
        tp = _p(trans, _f32p)
    lib().oracle_se3_transform(
        _p(quat, _f32p), tp, _p(pts, _f32p), ctypes.c_int64(n_parts),
        ctypes.c_int64(N), _p(out, _f32p))
    return out


def knn(x, k=20):
    """x [n, C, N] float32 -> sorted neighbour index sets [n, N, k] int64."""
    x = _f32(x)
    n, C, N = x.shape
    out = np.empty((n, N, k), np.int64)
    lib().oracle_knn(
        _p(x, _f32p), ctypes.c_int64(n), ctypes.c_int64(C), ctypes.c_int64(N),
        ctypes.c_int64(k), _p(out, _i64p))
    return out


def furthest_point_sample(xyz, m):
    """xyz [B, n, 3] float32 -> idx [B, m] int64 (pointnet2_ops furthest_point_sampling)."""
    xyz = _f32(xyz)
    B, n, _ = xyz.shape
    out = np.empty((B, m), np.int64)
    lib().oracle_fps(_p(xyz, _f32p), ctypes.c_int64(B), ctypes.c_int64(n), ctypes.c_int64(m), _p(out, _i64p))
    return out


def ball_query(xyz, new_xyz, radius, nsample):
    """idx [B, m, nsample] int64 (pointnet2_ops ball_query)."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, n, _ = xyz.shape
    m = new_xyz.shape[1]
    out = np.empty((B, m, nsample), np.int64)
    lib().oracle_ball_query(_p(xyz, _f32p), _p(new_xyz, _f32p), ctypes.c_int64(B), ctypes.c_int64(n),
                            ctypes.c_int64(m), ctypes.c_float(radius), ctypes.c_int64(nsample), _p(out, _i64p))
    return out
