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
torch.cuda.synchronize()
print('sanitize smoke done')
