import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200 import kernels
from multi_part_assembly_b200.models import build_encoder
dev = torch.device('cuda:0')
enc = build_encoder('pointnet', 256).to(dev).train()
x = torch.rand(640, 1000, 3, device=dev) - 0.5
kernels.set_precision('bf16')
for _ in range(2):
    with torch.no_grad():
        enc(x)
torch.cuda.synchronize()
