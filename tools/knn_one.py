import sys, torch
sys.path.insert(0, '/root/repo')
from multi_part_assembly_b200 import kernels
x = torch.randn(256, 1000, 64, device='cuda')
kernels.knn(x, 20); torch.cuda.synchronize()
