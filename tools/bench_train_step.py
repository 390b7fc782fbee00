"""Full training step (forward + loss + backward + optimizer) of pn_transformer cfg C on one
B200: native package vs the reference GPU build, eager, bf16 autocast on
both, CUDA events, median.   gpurun -- python tools/bench_train_step.py -> gpurun_out/train_step.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, 'baseline', '_ref')
sys.path.insert(0, REF)
import torch  # noqa: E402

HAVE_REF = os.path.isdir(os.path.join(REF, 'multi_part_assembly')) and \
    os.path.exists(os.path.join(REF, 'chamfer_cuda.so'))
if HAVE_REF:
    from oracle import ref_shims  # noqa: E402
    ref_shims.install(root=REF, cuda_chamfer=True)
    from multi_part_assembly.models import build_model as ref_build_model  # noqa: E402
from multi_part_assembly_b200.configs import get_cfg  # noqa: E402
from multi_part_assembly_b200.datasets import make_batch  # noqa: E402
from multi_part_assembly_b200.models import build_model  # noqa: E402
from multi_part_assembly_b200.compat.lightning import Trainer  # noqa: E402
from multi_part_assembly_b200 import profiler  # noqa: E402

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32


def prep(model):
    model = model.to(dev).train()
    model.trainer = Trainer()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
        if hasattr(m, 'dropout') and isinstance(m.dropout, float):
            m.dropout = 0.0
    return model


def measure(model, iters=10, warm=3):
    opt = model.configure_optimizers()
    if isinstance(opt, tuple):
        opt = opt[0][0]
    batch = make_batch(B, P=20, N=1000, num_valid=20, seed=0, device=dev)

    def step():
        with torch.autocast('cuda', dtype=torch.bfloat16):
            loss = model.training_step(dict(batch), 0)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warm):
        step()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); loss = step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], float(loss)


def measure_graphed(model, iters=20, warm=3):
    from multi_part_assembly_b200.runtime import GraphedTrainStep
    opt = model.configure_optimizers()
    if isinstance(opt, tuple):
        opt = opt[0][0]
    batch = make_batch(B, P=20, N=1000, num_valid=20, seed=0, device=dev)
    g = GraphedTrainStep(model, opt, batch)
    for _ in range(warm):
        g()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); loss = g(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], float(loss)


res = {'B': B}
cfg = get_cfg('pn_transformer', 'everyday')
torch.manual_seed(0)
ms, loss = measure(prep(build_model(cfg)))
res['native_train_step_ms'] = ms
res['native_train_shapes_per_s'] = B / ms * 1e3
print(f'native   train step (eager): {ms:.3f} ms  ({B / ms * 1e3:.0f} shapes/s)  loss {loss:.4f}', flush=True)
try:
    torch.manual_seed(0)
    ms, loss = measure_graphed(prep(build_model(cfg)))
    res['native_graph_train_step_ms'] = ms
    res['native_graph_train_shapes_per_s'] = B / ms * 1e3
    print(f'native   train step (CUDA graph): {ms:.3f} ms  ({B / ms * 1e3:.0f} shapes/s)  loss {loss:.4f}',
          flush=True)
except Exception as e:  # a host sync somewhere in the step
    res['native_graph_error'] = repr(e)[:400]
    print('graph capture failed:', repr(e)[:400], flush=True)
    torch.cuda.synchronize()
if HAVE_REF:
    ms, loss = measure(prep(ref_build_model(cfg)))
    res['reference_gpu_train_step_ms'] = ms
    res['reference_gpu_train_shapes_per_s'] = B / ms * 1e3
    best = max(res['native_train_shapes_per_s'], res.get('native_graph_train_shapes_per_s', 0.))
    res['speedup'] = best / res['reference_gpu_train_shapes_per_s']
    print(f'reference train step: {ms:.3f} ms  ({B / ms * 1e3:.0f} shapes/s)  loss {loss:.4f}', flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'train_step.json'), 'w'), indent=1)
print(json.dumps(res))
