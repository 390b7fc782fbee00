"""One eager training step inside a cudaProfilerStart/Stop range (for ncu -k regex:bn_ ...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200.configs import get_cfg
from multi_part_assembly_b200.datasets import make_batch
from multi_part_assembly_b200.models import build_model
from multi_part_assembly_b200.compat.lightning import Trainer
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = build_model(get_cfg('pn_transformer')).to(dev).train()
model.trainer = Trainer()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, 'dropout') and isinstance(m.dropout, float): m.dropout = 0.0
opt = model.configure_optimizers()
if isinstance(opt, tuple): opt = opt[0][0]
batch = make_batch(32, P=20, N=1000, num_valid=20, seed=0, device=dev)

def step():
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss = model.training_step(dict(batch), 0)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()

for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
