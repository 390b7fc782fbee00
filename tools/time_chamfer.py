"""Scratch timing of the Chamfer kernels through the C ABI with preallocated
buffers (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200 import _lib
from multi_part_assembly_b200.utils.transforms import random_quaternions

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


torch.manual_seed(0)
res = {}
only = sys.argv[1] if len(sys.argv) > 1 else ''
if only in ('', 'generic'):
    for (B, N) in [(640, 1000), (32, 20000), (32, 10240), (32, 40960)]:
        x1 = torch.rand(B, N, 3, device=dev) - 0.5
        x2 = torch.rand(B, N, 3, device=dev) - 0.5
        d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, N, device=dev)
        i1 = torch.empty(B, N, dtype=torch.int64, device=dev); i2 = torch.empty_like(i1)
        for name, algo in (('brute', 1), ('grid', 2)):
            wsb = L.mpa_chamfer_forward_workspace_bytes(B, N, N, algo)
            ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
            fn = lambda: L.mpa_chamfer_forward(x1.data_ptr(), x2.data_ptr(), B, N, N, d1.data_ptr(), i1.data_ptr(),
                                               d2.data_ptr(), i2.data_ptr(), algo, ws.data_ptr(), wsb, stream)
            ms = timeit(fn, iters=5 if algo == 1 else 20)
            res[f'chamfer_{name}_{B}x{N}'] = ms
            print(f'chamfer {name:5s} B={B} N={N}: {ms:.3f} ms  pairs/s={2*B*N*N/ms/1e9:.2f}e12 '
                  f'algGB/s={24*B*2*N/ms/1e6:.1f}', flush=True)

if only in ('', 'pose'):
    B, P, N = 32, 20, 1000
    pts = torch.rand(B, P, N, 3, device=dev) - 0.5
    pts = pts - pts.mean(2, keepdim=True)
    q1 = random_quaternions((B, P)).to(dev); q2 = random_quaternions((B, P)).to(dev)
    t1 = torch.randn(B, P, 3, device=dev) * 0.1; t2 = torch.rand(B, P, 3, device=dev) - 0.5
    valids = torch.ones(B, P, device=dev)
    d1 = torch.empty(B, P, N, device=dev); d2 = torch.empty_like(d1)
    i1 = torch.empty(B, P, N, dtype=torch.int32, device=dev); i2 = torch.empty_like(i1)
    p1 = torch.empty(B, P, N, 3, device=dev); p2 = torch.empty_like(p1)
    for mode in (0, 1):
        wsb = L.mpa_pose_chamfer_workspace_bytes(B, P, N, mode)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        fn = lambda: L.mpa_pose_chamfer(pts.data_ptr(), q1.data_ptr(), t1.data_ptr(), q2.data_ptr(), t2.data_ptr(),
                                        valids.data_ptr(), B, P, N, mode, d1.data_ptr(), i1.data_ptr(), d2.data_ptr(),
                                        i2.data_ptr(), p1.data_ptr(), p2.data_ptr(), ws.data_ptr(), wsb, stream)
        ms = timeit(fn, iters=20)
        res[f'pose_chamfer_mode{mode}'] = ms
        print(f'pose_chamfer mode={mode}: {ms:.3f} ms', flush=True)
print(json.dumps(res))
