"""Run the native transformer encoder forward a few times (B=32 shapes x 20 tokens, d=256,
bf16 mode) -- a small target for `ncu -k regex:encoder_ffn_block|linear_bf16|attention`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from multi_part_assembly_b200 import kernels, profiler  # noqa: E402
from multi_part_assembly_b200.models.pn_transformer import TransformerEncoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda:0')
torch.manual_seed(0)
tr = TransformerEncoder(256, 8, 1024, 4).to(dev).eval()
tokens = torch.randn(B, 20, 256, device=dev)
valid = torch.ones(B, 20, dtype=torch.bool, device=dev)
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
kernels.set_precision('bf16')
with torch.no_grad():
    for _ in range(3):
        tr(tokens, valid)
    torch.cuda.synchronize()
    profiler.enable(True)
    for _ in range(10):
        flush.zero_()
        tr(tokens, valid)
    torch.cuda.synchronize()
rep = profiler.report()
tot = 0.0
for k, v in sorted(rep.items()):
    print(f'{k:32s} {v["launches"]:4d} launches  {1e3 * v["ms_total"] / v["launches"]:8.2f} us each  '
          f'{1e3 * v["ms_total"] / 10:8.1f} us per forward')
    tot += v['ms_total'] / 10
print(f'total {1e3 * tot:.1f} us per forward')

# cycle stamps of the fused block (CTA 0, last launch): MPA_FFN_DEBUG = device pointer
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
os.environ['MPA_FFN_DEBUG'] = str(dbg.data_ptr())
with torch.no_grad():
    tr(tokens, valid)
torch.cuda.synchronize()
del os.environ['MPA_FFN_DEBUG']
d = dbg.cpu().tolist()
t0 = d[0]
rel = lambda i: (d[i] - t0) if d[i] else None
print('kernel entry -> MMA thread start (cycles):', d[0] - d[61], ' entry -> cluster reduce done (ns):', d[62] - d[60])
print('MMA thread: att_full', rel(1), 'out_proj issued', rel(2), 'a1_ready', rel(3), 'all issued', rel(4))
print('MMA thread hid_ready waits done at', [rel(8 + c) for c in range(8)])
print('epilogue: E1', rel(32), rel(33), '(residual slabs done', rel(37), ')  E3', rel(34), rel(35), ' cluster reduce done', rel(36))
print('epilogue E2 (start, end) per chunk', [(rel(40 + 2 * c), rel(41 + 2 * c)) for c in range(8)])
