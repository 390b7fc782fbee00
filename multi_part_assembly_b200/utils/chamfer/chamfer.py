"""Chamfer distance operator, same public API as the reference's
multi_part_assembly/utils/chamfer/chamfer.py (:11-76), backed by the sm_100a
kernels of csrc/chamfer.cu through the C ABI (include/mpa_b200.h)."""
import torch
from torch.amp import custom_bwd, custom_fwd

from ... import _lib

ALGO_AUTO, ALGO_BRUTE, ALGO_GRID = 0, 1, 2


def safe_sqrt(x, eps=1e-12):
    return torch.sqrt(torch.clamp(x, eps))


def chamfer_forward(xyz1, xyz2, algo=ALGO_AUTO, need_idx=True):
    """[B,N1,3],[B,N2,3] fp32 CUDA -> dist1, idx1 (int64), dist2, idx2.

    Mirrors chamfer_cuda.chamfer_forward (reference chamfer.cpp:21).
    """
    _lib.require_cuda(xyz1, xyz2)
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or \
            xyz2.shape[2] != 3 or xyz1.shape[0] != xyz2.shape[0]:
        raise RuntimeError(
            f'chamfer_forward expects (B,N1,3) and (B,N2,3), got '
            f'{tuple(xyz1.shape)} and {tuple(xyz2.shape)}')
    if xyz1.dtype != torch.float32 or xyz2.dtype != torch.float32:
        raise RuntimeError('chamfer_forward expects float32 inputs')
    xyz1 = xyz1.contiguous()
    xyz2 = xyz2.contiguous()
    B, n1, _ = xyz1.shape
    n2 = xyz2.shape[1]
    dev = xyz1.device
    # the reference returns zero-initialised outputs (chamfer_kernel.cu:129-132)
    dist1 = torch.zeros(B, n1, dtype=torch.float32, device=dev)
    dist2 = torch.zeros(B, n2, dtype=torch.float32, device=dev)
    idx1 = torch.zeros(B, n1, dtype=torch.int64, device=dev) if need_idx else None
    idx2 = torch.zeros(B, n2, dtype=torch.int64, device=dev) if need_idx else None
    L = _lib.lib()
    ws_bytes = L.mpa_chamfer_forward_workspace_bytes(B, n1, n2, algo)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    with torch.cuda.device(dev):
        rc = L.mpa_chamfer_forward(
            _lib.ptr(xyz1), _lib.ptr(xyz2), B, n1, n2, _lib.ptr(dist1),
            _lib.ptr(idx1), _lib.ptr(dist2), _lib.ptr(idx2), algo,
            _lib.ptr(ws), ws_bytes, _lib.cuda_stream(dev))
    _lib.check(rc, 'mpa_chamfer_forward')
    return dist1, idx1, dist2, idx2


def chamfer_backward(grad_dist1, grad_dist2, xyz1, xyz2, idx1, idx2):
    """Mirrors chamfer_cuda.chamfer_backward (reference chamfer.cpp:22)."""
    _lib.require_cuda(grad_dist1, grad_dist2, xyz1, xyz2, idx1, idx2)
    B, n1, _ = xyz1.shape
    n2 = xyz2.shape[1]
    dev = xyz1.device
    grad_xyz1 = torch.empty(B, n1, 3, dtype=torch.float32, device=dev)
    grad_xyz2 = torch.empty(B, n2, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().mpa_chamfer_backward(
            _lib.ptr(grad_dist1), _lib.ptr(grad_dist2), _lib.ptr(xyz1),
            _lib.ptr(xyz2), _lib.ptr(idx1), _lib.ptr(idx2), B, n1, n2,
            _lib.ptr(grad_xyz1), _lib.ptr(grad_xyz2), _lib.cuda_stream(dev))
    _lib.check(rc, 'mpa_chamfer_backward')
    return grad_xyz1, grad_xyz2


class ChamferDistanceFunction(torch.autograd.Function):
    """Same contract as the reference Function (chamfer.py:11-33): inputs are
    force-cast to float32 under autocast, indices are saved for backward."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, xyz1, xyz2):
        if xyz1.dtype == torch.float64 or xyz2.dtype == torch.float64:
            raise RuntimeError(
                'chamfer_distance: float64 is not supported by the sm_100a '
                'kernels (the reference only uses it in its gradcheck test)')
        ctx.in_dtypes = (xyz1.dtype, xyz2.dtype)
        # the kernels read fp32: cast ONCE and save the cast tensors, so that the
        # backward never hands half-precision buffers to a float* entry point
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        dist1, idx1, dist2, idx2 = chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        grad_dist1 = grad_dist1.contiguous().float()
        grad_dist2 = grad_dist2.contiguous().float()
        g1, g2 = chamfer_backward(grad_dist1, grad_dist2, xyz1, xyz2, idx1, idx2)
        return g1.to(ctx.in_dtypes[0]), g2.to(ctx.in_dtypes[1])


def chamfer_distance(xyz1, xyz2, transpose=False, sqrt=False, eps=1e-12):
    """Chamfer distance, reference signature (chamfer.py:36-64).

    Args:
        xyz1: (b, n1, 3) or (n1, 3); xyz2: (b, n2, 3) or (n2, 3)
        transpose: inputs are (b, 3, n)
        sqrt: return sqrt of the squared distances (clamped at eps)
    Returns:
        dist1 (b, n1), dist2 (b, n2)
    """
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1 = xyz1.transpose(1, 2)
        xyz2 = xyz2.transpose(1, 2)
    dist1, dist2 = ChamferDistanceFunction.apply(xyz1, xyz2)
    if sqrt:
        dist1 = safe_sqrt(dist1, eps)
        dist2 = safe_sqrt(dist2, eps)
    return dist1, dist2


def nn_distance(xyz1, xyz2, transpose=True):
    """Inference interface (reference chamfer.py:67-76): dist1, idx1, dist2, idx2."""
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1 = xyz1.transpose(1, 2).contiguous()
        xyz2 = xyz2.transpose(1, 2).contiguous()
    return chamfer_forward(xyz1, xyz2)
