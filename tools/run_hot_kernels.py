"""Scratch: run a few eager bench steps at cfg C sizes (for `ncu --set full` captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200.configs import get_cfg
from multi_part_assembly_b200.datasets import make_batch
from multi_part_assembly_b200.models import build_model
from multi_part_assembly_b200.compat.lightning import Trainer
dev = torch.device('cuda:0')
model = build_model(get_cfg('pn_transformer')).to(dev).train()
model.trainer = Trainer()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, 'dropout') and isinstance(m.dropout, float): m.dropout = 0.0
batch = make_batch(32, P=20, N=1000, num_valid=20, seed=0, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        loss = model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']
torch.cuda.synchronize()
print(float(loss))
