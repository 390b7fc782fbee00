"""One CUDA-graph replay of the bench step inside a cudaProfilerStart/Stop range
(use with `ncu --profile-from-start off --graph-profiling node`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200.configs import get_cfg
from multi_part_assembly_b200.datasets import make_batch
from multi_part_assembly_b200.models import build_model
from multi_part_assembly_b200.compat.lightning import Trainer
from multi_part_assembly_b200.runtime import GraphedStep
dev = torch.device('cuda:0')
torch.manual_seed(0)  # same weights (hence predicted poses) as bench.py rank 0
model = build_model(get_cfg('pn_transformer')).to(dev).train()
model.trainer = Trainer()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, 'dropout') and isinstance(m.dropout, float): m.dropout = 0.0
batch = make_batch(32, P=20, N=1000, num_valid=20, seed=0, device=dev)
g = GraphedStep(model, batch)
for _ in range(3):
    g()
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = g()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(float(out['loss']))
