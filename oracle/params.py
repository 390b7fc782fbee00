"""Deterministic parameter filling so golden fixtures need not store weights
(TEST INFRASTRUCTURE).  Same numpy stream on every platform: parameters and
buffers are visited in sorted-name order."""
import numpy as np
import torch


@torch.no_grad()
def fill_params_(module, seed):
    rng = np.random.default_rng(seed)
    tensors = dict(module.named_parameters())
    tensors.update(dict(module.named_buffers()))
    done = set()
    for name in sorted(tensors):
        t = tensors[name]
        if id(t) in done or not t.dtype.is_floating_point:
            continue
        done.add(id(t))
        leaf = name.split('.')[-1]
        if leaf == 'running_var':
            v = rng.uniform(0.5, 1.5, t.shape)
        elif leaf == 'running_mean':
            v = rng.normal(0, 0.1, t.shape)
        elif t.dim() == 1 and leaf == 'weight':  # norm scales
            v = rng.uniform(0.5, 1.5, t.shape) * rng.choice([-1.0, 1.0], t.shape, p=[0.1, 0.9])
        elif t.dim() == 1:  # biases
            v = rng.normal(0, 0.05, t.shape)
        else:
            fan_in = int(np.prod(t.shape[1:]))
            v = rng.normal(0, 1.0 / np.sqrt(fan_in), t.shape)
        t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)))
    return module


def zero_dropout(module):
    """Parity runs compare with dropout disabled on both sides."""
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
        if hasattr(m, 'dropout') and isinstance(getattr(m, 'dropout'), float):
            m.dropout = 0.0
    return module
