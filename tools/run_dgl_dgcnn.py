"""Scratch: BASELINE config D (configs/dgl + DGCNN encoder, 16 valid parts of 1000 points,
B=32) forward + loss, eager, timed with CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_part_assembly_b200.configs import get_cfg
from multi_part_assembly_b200.datasets import make_batch
from multi_part_assembly_b200.models import build_model
from multi_part_assembly_b200.compat.lightning import Trainer
from multi_part_assembly_b200 import profiler
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
enc = sys.argv[3] if len(sys.argv) > 3 else 'dgcnn'
cfg = get_cfg('dgl', 'everyday', encoder=enc)
model = build_model(cfg).to(dev).train()
model.trainer = Trainer()
batch = make_batch(B, P=20, N=1000, num_valid=16, seed=0, device=dev)

def step():
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        return model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']

for _ in range(2):
    loss = step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    loss = step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(f'dgl+dgcnn cfg D: B={B}  {ms:.2f} ms/step  {B / ms * 1e3:.0f} shapes/s  loss={float(loss):.4f}  '
      f'peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB')
profiler.enable(True)
step(); torch.cuda.synchronize()
for k, v in sorted(profiler.report().items(), key=lambda kv: -kv[1]['ms_total']):
    print(f'   {k:32s} {v["launches"]:4d} launches {v["ms_total"]:8.3f} ms')
