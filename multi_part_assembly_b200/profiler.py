"""Per-kernel CUDA-event timing of the library's own launches (bench.py)."""
import ctypes

from . import _lib


def enable(on=True):
    _lib.lib().mpa_profile_enable(1 if on else 0)


def report():
    """{kernel: {'launches': n, 'ms_total': t}} since the last report."""
    L = _lib.lib()
    buf = ctypes.create_string_buffer(1 << 16)
    L.mpa_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split()
        out[name] = {'launches': int(n), 'ms_total': float(ms)}
    return out


def pair_stats(step):
    """Candidate pairs the exact Chamfer searches evaluate in ONE call of `step()` (a callable
    running one eager forward+loss), keyed like the kernel names of `report()`."""
    import torch
    L = _lib.lib()
    out = (ctypes.c_ulonglong * 3)()
    _lib.check(L.mpa_chamfer_pair_count(1, out), 'mpa_chamfer_pair_count')  # on + clear
    try:
        step()
        torch.cuda.synchronize()
    finally:
        _lib.check(L.mpa_chamfer_pair_count(0, out), 'mpa_chamfer_pair_count')  # off + read
    return {'chamfer_grid_nn_part': float(out[0]), 'chamfer_grid_nn_shape': float(out[1]),
            'chamfer_grid_nn': float(out[2])}
