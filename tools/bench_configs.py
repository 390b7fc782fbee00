"""BASELINE.json configs B-E on one B200: native (CUDA-graph replay, bf16 tensor-core
GEMMs) next to the "reference GPU build" (unmodified reference Python +
its own Chamfer kernels for sm_100a + stock torch), same synthetic batches, forward + loss
without autograd, CUDA events, L2 flushed, median of `iters`.

    gpurun -- python tools/bench_configs.py    -> gpurun_out/configs.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, 'baseline', '_ref')
sys.path.insert(0, REF)

import torch  # noqa: E402

HAVE_REF = os.path.exists(os.path.join(REF, 'chamfer_cuda.so')) and \
    os.path.isdir(os.path.join(REF, 'multi_part_assembly'))
if HAVE_REF:
    from oracle import ref_shims  # noqa: E402
    ref_shims.install(root=REF, cuda_chamfer=True)
    from multi_part_assembly.models import build_model as ref_build_model  # noqa: E402
from multi_part_assembly_b200.configs import get_cfg  # noqa: E402
from multi_part_assembly_b200.datasets import make_batch  # noqa: E402
from multi_part_assembly_b200.models import build_model  # noqa: E402
from multi_part_assembly_b200.compat.lightning import Trainer  # noqa: E402
from multi_part_assembly_b200.runtime import GraphedStep  # noqa: E402

dev = torch.device('cuda:0')
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
        if hasattr(m, 'dropout') and isinstance(m.dropout, float):
            m.dropout = 0.0
    return model


CONFIGS = [  # tag, model, cfg kwargs, B, valid parts, N
    ('B_global', 'global', {}, 32, 8, 1000),
    ('C_pn_transformer', 'pn_transformer', {}, 32, 20, 1000),
    ('D_dgl_dgcnn', 'dgl', {'encoder': 'dgcnn'}, 32, 16, 1000),
    ('D_dgl_pointnet', 'dgl', {}, 32, 16, 1000),
    ('E_pn_transformer_N512', 'pn_transformer', {}, 32, 20, 512),
    ('E_pn_transformer_N2048', 'pn_transformer', {}, 32, 20, 2048),
    ('E_pn_transformer_B256', 'pn_transformer', {}, 256, 20, 1000),
]
only = sys.argv[1:]
res = {}
for tag, name, kw, B, nv, N in CONFIGS:
    if only and tag not in only:
        continue
    cfg = get_cfg(name, 'everyday', **kw)
    batch = make_batch(B, P=20, N=N, num_valid=nv, seed=0, device=dev)
    entry = {'B': B, 'valid_parts': nv, 'N': N}
    try:
        model = no_dropout(build_model(cfg)).to(dev).train()
        model.trainer = Trainer()

        def eager():
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                return model.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']

        ms = timeit(eager)
        entry['native_eager_ms'] = ms
        try:
            g = GraphedStep(model, batch)
            ms_g = timeit(lambda: g())
            entry['native_graph_ms'] = ms_g
            ms = min(ms, ms_g)
            del g
        except Exception as e:  # e.g. a host sync inside the step
            entry['native_graph_error'] = repr(e)[:200]
        entry['native_shapes_per_s'] = B / ms * 1e3
        entry['native_peak_mem_gib'] = torch.cuda.max_memory_allocated() / 2**30
        del model
    except Exception as e:
        entry['native_error'] = repr(e)[:300]
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    if HAVE_REF:
        try:
            t0 = time.time()
            rmodel = no_dropout(ref_build_model(cfg)).to(dev).train()
            rmodel.trainer = Trainer()

            def rstep():
                with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
                    return rmodel.forward_pass(dict(batch), mode='train', optimizer_idx=-1)['loss']

            ms_r = timeit(rstep, iters=5, warm=2)
            entry['reference_gpu_fp16_ms'] = ms_r
            entry['reference_gpu_shapes_per_s'] = B / ms_r * 1e3
            entry['reference_peak_mem_gib'] = torch.cuda.max_memory_allocated() / 2**30
            if 'native_shapes_per_s' in entry:
                entry['speedup'] = entry['native_shapes_per_s'] / entry['reference_gpu_shapes_per_s']
            del rmodel
        except Exception as e:
            entry['reference_error'] = repr(e)[:300]
        torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    res[tag] = entry
    print(tag, json.dumps(entry), flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'configs.json'), 'w'), indent=1)
