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
