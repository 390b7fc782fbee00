"""Measure the "reference GPU build"  on the same B200:
the UNMODIFIED reference Python (pip-installed copy in baseline/_ref) + its own
Chamfer CUDA kernels compiled for sm_100a (baseline/build_ref_chamfer.py) + stock
torch 2.11 kernels, on the same synthetic batch and with the same timing
harness as bench.py (CUDA events, L2 flushed, dropout 0, no autograd recording).

    gpurun -- python tools/bench_reference_gpu.py        -> gpurun_out/reference_gpu.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, 'baseline', '_ref')
sys.path.insert(0, REF)

import torch  # noqa: E402
from oracle import ref_shims  # noqa: E402

ref_shims.install(root=REF, cuda_chamfer=True)
from multi_part_assembly.models import build_model  # noqa: E402
from multi_part_assembly.utils import chamfer_distance  # noqa: E402
from multi_part_assembly_b200.configs import get_cfg  # noqa: E402
from multi_part_assembly_b200.datasets import make_batch  # noqa: E402
from multi_part_assembly_b200.compat.lightning import Trainer  # noqa: E402

dev = torch.device('cuda:0')
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
torch.manual_seed(0)
for (B, N) in [(640, 1000), (32, 20000)]:
    x1 = torch.rand(B, N, 3, device=dev) - 0.5
    x2 = torch.rand(B, N, 3, device=dev) - 0.5
    ms = timeit(lambda: chamfer_distance(x1, x2), 10)
    res[f'ref_chamfer_{B}x{N}_ms'] = ms
    print(f'reference chamfer_distance [{B},{N},3]^2: {ms:.3f} ms', flush=True)

model = build_model(get_cfg('pn_transformer')).to(dev).train()
model.trainer = Trainer()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = 0.0
    if isinstance(m, torch.nn.MultiheadAttention):
        m.dropout = 0.0
batch = make_batch(32, P=20, N=1000, num_valid=20, seed=0, device=dev)
for name, amp, grad in (('fp32_nograd', None, False), ('fp16_nograd', torch.float16, False),
                        ('bf16_nograd', torch.bfloat16, False), ('fp16_grad', torch.float16, True)):

    def step():
        with torch.set_grad_enabled(grad):
            with torch.autocast('cuda', dtype=amp or torch.float16, enabled=amp is not None):
                return model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']

    ms = timeit(step, 10)
    res[f'ref_pn_transformer_fwd_loss_{name}_ms'] = ms
    res[f'ref_pn_transformer_fwd_loss_{name}_shapes_per_s'] = 32 / ms * 1e3
    print(f'reference pn_transformer fwd+loss {name}: {ms:.3f} ms = {32 / ms * 1e3:.0f} shapes/s', flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'reference_gpu.json'), 'w'), indent=1)
print(json.dumps(res))
