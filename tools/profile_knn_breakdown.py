"""Per-kernel CUDA-event times of one k-NN call (cfg D layer shapes), tensor-core path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200 import kernels, profiler
dev = torch.device('cuda:0')
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for C in (3, 64, 128):
    x = torch.randn(n, 1000, C, device=dev)
    for _ in range(2):
        kernels.knn(x, 20)
    torch.cuda.synchronize()
    profiler.enable(True)
    kernels.knn(x, 20)
    torch.cuda.synchronize()
    rep = profiler.report()
    profiler.enable(False)
    print(f'C={C}: ' + ', '.join(f'{k} {v["launches"]}x {1e3 * v["ms_total"] / v["launches"]:.1f} us = {v["ms_total"]:.3f} ms'
                                 for k, v in sorted(rep.items())))
