"""Scratch timing of the Chamfer kernels (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200.utils.chamfer import chamfer_forward
from multi_part_assembly_b200.utils.loss import pose_chamfer
from multi_part_assembly_b200.utils.transforms import random_quaternions

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


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
for (B, N) in [(640, 1000), (32, 20000), (32, 10240), (32, 40960)]:
    x1 = torch.rand(B, N, 3, device=dev) - 0.5
    x2 = torch.rand(B, N, 3, device=dev) - 0.5
    for name, algo in (('brute', 1), ('grid', 2)):
        ms = timeit(lambda: chamfer_forward(x1, x2, algo=algo), iters=5 if algo == 1 else 20)
        res[f'chamfer_{name}_{B}x{N}'] = ms
        print(f'chamfer {name:5s} B={B} N={N}: {ms:.3f} ms  pair-evals/s={2*B*N*N/ms/1e9:.1f} G', flush=True)

B, P, N = 32, 20, 1000
pts = torch.rand(B, P, N, 3, device=dev) - 0.5
pts = pts - pts.mean(2, keepdim=True)
q1 = random_quaternions((B, P)).to(dev); q2 = random_quaternions((B, P)).to(dev)
t1 = torch.randn(B, P, 3, device=dev) * 0.1; t2 = torch.rand(B, P, 3, device=dev) - 0.5
valids = torch.ones(B, P, device=dev)
for mode in (0, 1):
    ms = timeit(lambda: pose_chamfer(pts, t1, t2, q1, q2, valids, mode), iters=20)
    res[f'pose_chamfer_mode{mode}'] = ms
    print(f'pose_chamfer mode={mode}: {ms:.3f} ms', flush=True)
print(json.dumps(res))
