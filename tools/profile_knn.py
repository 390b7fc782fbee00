"""Scratch: time the k-NN kernel alone (cfg D layer shapes) and give ncu one launch to look at."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200 import kernels
dev = torch.device('cuda:0')
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for C in (3, 64, 128):
    x = torch.randn(n, 1000, C, device=dev)
    for _ in range(2):
        kernels.knn(x, 20)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        kernels.knn(x, 20)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    fma = n * 1000 * 1000 * C
    print(f'knn n={n} N=1000 C={C}: {ms:.3f} ms  {fma / ms / 1e9:.2f} TFMA/s (peak 36.4: 148 SM x 128 lanes x 1.92 GHz)')
