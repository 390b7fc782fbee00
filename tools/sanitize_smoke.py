"""Small invocations of every native kernel (for compute-sanitizer memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200 import kernels
from multi_part_assembly_b200.configs import get_cfg
from multi_part_assembly_b200.datasets import make_batch
from multi_part_assembly_b200.models import build_model
from multi_part_assembly_b200.compat.lightning import Trainer
from multi_part_assembly_b200.utils.chamfer import chamfer_forward
dev = torch.device('cuda:0')
torch.manual_seed(0)
# chamfer: brute + grid, ragged
for n1, n2, algo in ((100, 37, 1), (700, 650, 2), (1000, 1000, 0)):
    chamfer_forward(torch.rand(2, n1, 3, device=dev), torch.rand(2, n2, 3, device=dev), algo=algo)
# whole models, no grad (native encoder / transformer / pose head / fused losses) and with grad
for name, enc in (('pn_transformer', 'pointnet'), ('dgl', 'dgcnn')):
    model = build_model(get_cfg(name, encoder=enc)).to(dev).train()
    model.trainer = Trainer()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
        if hasattr(m, 'dropout') and isinstance(m.dropout, float): m.dropout = 0.0
    batch = make_batch(2, P=20, N=200, num_valid=[5, 3], seed=0, device=dev)
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        print(name, float(model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']))
    loss = model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']
    loss.backward()
# reference training configuration: dropout 0.1 inside the native transformer (bf16) and the
# fp32-accurate three-plane mode
model = build_model(get_cfg('pn_transformer')).to(dev).train()
model.trainer = Trainer()
batch = make_batch(2, P=20, N=200, num_valid=[20, 7], seed=1, device=dev)
with torch.autocast('cuda', dtype=torch.bfloat16):
    model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss'].backward()
kernels.set_precision('fp32')
with torch.no_grad():
    print('fp32 mode', float(model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']))
kernels.set_precision('auto')
# PointNet++ ops
xyz = torch.rand(3, 500, 3, device=dev)
fi, cen = kernels.furthest_point_sample(xyz, 64)
kernels.ball_query(0.2, 16, xyz, cen)
for enc in ('pointnet2_ssg', 'pointnet2_msg'):
    m2 = build_model(get_cfg('pn_transformer', encoder=enc)).to(dev).train()
    m2.trainer = Trainer()
    with torch.no_grad():
        print(enc, float(m2.forward_pass(dict(make_batch(2, P=20, N=256, num_valid=[4, 3], seed=2, device=dev)),
                                         mode='train', optimizer_idx=-1)['loss']))
torch.cuda.synchronize()
print('sanitize smoke done')
