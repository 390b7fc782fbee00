"""Small helpers kept from the reference's utils/utils.py that the hot-path
models use (`_get_clones` :128, `filter_wd_parameters` :90-125)."""
import copy

import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm
from torch.nn.modules.instancenorm import _InstanceNorm


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


def filter_wd_parameters(model, skip_list=()):
    """Split parameters into weight-decayed and not (norm weights and all
    biases are not decayed), in sorted-name order."""
    norm_types = (nn.LayerNorm, nn.GroupNorm, _BatchNorm, _InstanceNorm)
    no_decay, seen = [], set()

    def add(p):
        if p is not None and id(p) not in seen:
            seen.add(id(p))
            no_decay.append(p)

    mods = dict(model.named_modules())
    for name in sorted(n for n, m in mods.items() if isinstance(m, norm_types)):
        add(mods[name].weight)
    for name in sorted(n for n, m in mods.items()
                       if getattr(m, 'bias', None) is not None):
        add(mods[name].bias)
    for name in sorted(n for n in mods if n in skip_list):
        for p in mods[name].parameters():
            if p.requires_grad:
                add(p)
    decay = [p for n, p in sorted(model.named_parameters())
             if p.requires_grad and id(p) not in seen]
    return {'decay': decay, 'no_decay': no_decay}
