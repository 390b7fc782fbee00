import os, sys, torch
sys.path.insert(0, '/root/repo')
from multi_part_assembly_b200 import kernels, profiler
from multi_part_assembly_b200.models.modules.encoder import build_encoder
torch.manual_seed(0)
dev = torch.device('cuda:0')
enc = build_encoder('pointnet', 256).to(dev).train()
x = torch.randn(640, 1000, 3, device=dev) * 0.3
valids = torch.ones(640, device=dev); valids[17] = 0; valids[600:] = 0
kernels.set_precision('bf16')
with torch.no_grad():
    out = enc(x, valids=valids) if getattr(enc, 'supports_valids', False) else enc(x)
    rm = [b.running_mean.clone() for b in enc.modules() if isinstance(b, torch.nn.BatchNorm1d)]
    torch.save({'out': out.cpu(), 'rm': [r.cpu() for r in rm]}, sys.argv[1])
    for _ in range(3): enc(x, valids=valids)
    torch.cuda.synchronize(); profiler.enable(True)
    for _ in range(10): enc(x, valids=valids)
    torch.cuda.synchronize()
rep = profiler.report(); tot = 0
for k, v in sorted(rep.items()):
    print(f'{k:28s} {1e3*v["ms_total"]/v["launches"]:8.2f} us'); tot += v['ms_total']/10
print('total us per forward', 1e3*tot)
