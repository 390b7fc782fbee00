"""Time the two fused pose-Chamfer calls of the bench step in isolation (no stream overlap) on the
poses an untrained pn_transformer actually predicts for the bench batch (B=32, P=20, N=1000):
per-kernel CUDA events from the library profiler, L2 flushed between iterations.

    gpurun -- python tools/chamfer_bench_like.py [iters]      (ncu-friendly: few launches)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multi_part_assembly_b200 import _lib, profiler  # noqa: E402
from multi_part_assembly_b200.configs import get_cfg  # noqa: E402
from multi_part_assembly_b200.datasets import make_batch  # noqa: E402
from multi_part_assembly_b200.models import build_model  # noqa: E402
from multi_part_assembly_b200.utils import Rotation3D  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = build_model(get_cfg('pn_transformer', 'everyday')).to(dev).train()
batch = make_batch(32, P=20, N=1000, num_valid=20, seed=0, device=dev)
with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
    out = model.forward({k: batch[k] for k in ('part_pcs', 'part_valids', 'part_label', 'instance_label')})
q1 = out['rot'].rot.float().contiguous()
t1 = out['trans'].float().contiguous()
q2 = Rotation3D(batch['part_quat']).rot.float().contiguous()
t2 = batch['part_trans'].float().contiguous()
pts, valids = batch['part_pcs'].contiguous(), batch['part_valids'].contiguous()
B, P, N = 32, 20, 1000
L = _lib.lib()
d1 = torch.empty(B, P, N, device=dev); d2 = torch.empty_like(d1)
i1 = torch.empty(B, P, N, dtype=torch.int32, device=dev); i2 = torch.empty_like(i1)
p1 = torch.empty(B, P, N, 3, device=dev); p2 = torch.empty_like(p1)
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
ws = {m: torch.empty(L.mpa_pose_chamfer_workspace_bytes(B, P, N, m), dtype=torch.uint8, device=dev) for m in (0, 1)}


def call(mode):
    rc = L.mpa_pose_chamfer(pts.data_ptr(), q1.data_ptr(), t1.data_ptr() if mode else None, q2.data_ptr(),
                            t2.data_ptr() if mode else None, valids.data_ptr(), B, P, N, mode, d1.data_ptr(),
                            i1.data_ptr(), d2.data_ptr(), i2.data_ptr(), p1.data_ptr(), p2.data_ptr(),
                            ws[mode].data_ptr(), ws[mode].numel(), stream)
    _lib.check(rc, 'mpa_pose_chamfer')


for m in (0, 1):
    call(m)
torch.cuda.synchronize()
profiler.enable(True)
for _ in range(iters):
    for m in (0, 1):
        flush.zero_()
        call(m)
torch.cuda.synchronize()
rep = profiler.report()
profiler.enable(False)
res = {k: v['ms_total'] / v['launches'] for k, v in rep.items()}
res['pairs'] = profiler.pair_stats(lambda: (call(0), call(1)))
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
tag = os.environ.get('MPA_TAG', 'run')
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', f'chamfer_bench_like_{tag}.json'), 'w'), indent=1)
