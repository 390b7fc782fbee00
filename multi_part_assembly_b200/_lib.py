"""ctypes binding of csrc/libmpa_b200.so (the C ABI declared in include/mpa_b200.h).

The CUDA library is the ONLY compute path of this package: if it cannot be
loaded, or a call fails, a RuntimeError is raised -- there is no CPU or eager
PyTorch fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get('MPA_B200_LIB') or os.path.join(_HERE, 'csrc', 'libmpa_b200.so')  # override: kernel experiments
_LIB = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_size_t = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/mpa_b200.h
_SIGNATURES = {
    'mpa_last_error': (ctypes.c_char_p, []),
    'mpa_version': (c_int, []),
    'mpa_launch_count': (ctypes.c_uint64, []),
    'mpa_profile_enable': (None, [c_int]),
    'mpa_profile_report': (c_size_t, [ctypes.c_char_p, c_size_t]),
    'mpa_chamfer_pair_count': (c_int, [c_int, c_void_p]),
    'mpa_chamfer_forward_workspace_bytes': (c_size_t, [c_int] * 4),
    'mpa_chamfer_forward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_void_p, c_size_t, c_void_p]),
    'mpa_chamfer_backward': (c_int, [c_void_p] * 6 + [c_int] * 3 + [c_void_p] * 3),
    'mpa_chamfer_forward_host': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_void_p]),
    'mpa_se3_transform': (c_int, [c_void_p] * 3 + [c_int] * 2 + [c_void_p] * 2),
    'mpa_se3_transform_backward': (c_int, [c_void_p] * 3 + [c_int] * 2 + [c_void_p] * 4),
    'mpa_quat_fix_zero': (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p]),
    'mpa_pose_chamfer_workspace_bytes': (c_size_t, [c_int] * 4),
    'mpa_pose_chamfer': (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_void_p] * 7 +
                         [c_size_t, c_void_p]),
    'mpa_geometric_losses_workspace_bytes': (c_size_t, [c_int] * 2),
    'mpa_geometric_losses': (c_int, [c_void_p] * 10 + [c_int] * 5 + [c_void_p] * 3 + [c_size_t, c_void_p]),
    'mpa_pointnet_workspace_bytes': (c_size_t, [c_int]),
    'mpa_pointnet_workspace_bytes_n': (c_size_t, [c_int, c_int]),
    'mpa_pointnet_forward': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 5 +
                             [c_int, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p, c_size_t,
                              c_void_p]),
    'mpa_pointnet_forward_ex': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 5 +
                                [c_int, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    'mpa_pose_outputs': (c_int, [c_void_p, c_int, c_int] + [c_void_p] * 4 + [c_int] + [c_void_p] * 3),
    'mpa_pose_head_forward': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_int] + [c_void_p] * 4 + [c_int] + [c_void_p] * 3),
    'mpa_bn_stats': (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'mpa_bn_finalize': (c_int, [c_void_p, c_int, c_int, c_int] + [c_void_p] * 3 + [ctypes.c_float] +
                        [c_void_p] * 6),
    'mpa_bn_act': (c_int, [c_void_p] * 3 + [c_int, ctypes.c_longlong, c_int, c_int] + [c_void_p] * 3),
    'mpa_bn_backward': (c_int, [c_void_p] * 9 + [ctypes.c_longlong, c_int, c_int] + [c_void_p] * 4),
    'mpa_pool_argmax': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'mpa_lsap_batched': (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_void_p]),
    'mpa_match_parts_workspace_bytes': (c_size_t, [c_int] * 2),
    'mpa_match_parts': (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] + [c_int] * 4 + [c_void_p] * 5 +
                        [c_size_t, c_void_p]),
    'mpa_linear_workspace_bytes': (c_size_t, [c_int] * 3),
    'mpa_linear_workspace_bytes_ex': (c_size_t, [c_int] * 4),
    'mpa_linear_forward_ex': (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    'mpa_linear_forward': (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    'mpa_transformer_workspace_bytes': (c_size_t, [c_int] * 5),
    'mpa_transformer_mask_bytes': (c_size_t, [c_int] * 6),
    'mpa_transformer_forward': (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p] * 14 +
                                [ctypes.c_float, ctypes.c_float, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    'mpa_knn_workspace_bytes': (c_size_t, [c_int] * 2),
    'mpa_knn_workspace_bytes_c': (c_size_t, [c_int] * 3),
    'mpa_knn': (c_int, [c_void_p, c_void_p] + [c_int] * 4 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    'mpa_edge_aggregate_workspace_bytes': (c_size_t, [ctypes.c_longlong, c_int]),
    'mpa_edge_aggregate': (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p] * 4 +
                           [c_size_t, c_void_p]),
    'mpa_edgeconv_finish': (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p] * 4 +
                            [c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p,
                             c_int, c_int, c_void_p]),
    'mpa_bn_pool_workspace_bytes': (c_size_t, [c_int] * 2),
    'mpa_bn_pool': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 4 +
                    [c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p, c_size_t,
                     c_void_p]),
    'mpa_furthest_point_sample': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'mpa_ball_query': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_float, c_int, c_void_p,
                               c_void_p]),
    'mpa_group_rows': (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_void_p, c_void_p]),
    'mpa_column_stats_workspace_bytes': (c_size_t, [c_int] * 2),
    'mpa_column_stats': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'mpa_pose_chamfer_backward_workspace_bytes': (c_size_t, [c_int] * 3),
    'mpa_pose_chamfer_backward': (c_int, [c_void_p] * 10 + [c_int] * 4 +
                                  [c_void_p] * 5 + [c_size_t, c_void_p]),
}


def library_path():
    return _SO


def lib():
    """Load libmpa_b200.so (building it first if the sources are newer)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if 'MPA_B200_LIB' not in os.environ:
        # build in-tree when the library is missing or older than its sources (mtime check
        # and file lock inside build()); raises if compilation fails.  A box without nvcc
        # (the GPU box) loads the prebuilt library as it is.
        from .csrc.build import build, NVCC
        if os.path.exists(NVCC) or not os.path.exists(_SO):
            build()
    try:
        handle = ctypes.CDLL(_SO)
    except OSError as e:  # pragma: no cover
        raise RuntimeError(
            f'multi_part_assembly_b200: cannot load the CUDA library {_SO}: {e}. '
            'There is no CPU fallback; run `python -c "import __graft_entry__ as g; '
            'g.build()"` on a machine with nvcc.') from e
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _LIB = handle
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().mpa_last_error().decode(errors='replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def cuda_stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                'multi_part_assembly_b200 kernels need CUDA tensors (no CPU path); '
                f'got a tensor on {t.device}')


def launch_count():
    return int(lib().mpa_launch_count())
