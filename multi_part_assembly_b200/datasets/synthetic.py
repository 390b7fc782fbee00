"""Synthetic batches with the dataset dict schema of the reference
(datasets/geometry_data.py:173-207, partnet_data.py:146-237; SURVEY.md
appendix B and 8d "Synthetic inputs")."""
import torch

from ..utils.rotation_conversions import random_quaternions


def make_batch(B, P=20, N=1000, num_valid=20, seed=0, semantic=False, device='cpu',
               valid_matrix=True, pin_memory=False):
    """part_pcs ~ U(-0.5, 0.5)^3 recentred per part, part_trans ~ U(-0.5, 0.5)^3,
    unit scalar-first quaternions with w >= 0, the first `num_valid` (int or
    per-shape list) slots valid, everything padded with zeros, float32."""
    g = torch.Generator().manual_seed(seed)
    nv = [num_valid] * B if isinstance(num_valid, int) else list(num_valid)
    assert len(nv) == B and all(1 <= v <= P for v in nv)
    valids = torch.zeros(B, P)
    for b, v in enumerate(nv):
        valids[b, :v] = 1.
    pcs = torch.rand(B, P, N, 3, generator=g) - 0.5
    pcs = pcs - pcs.mean(dim=2, keepdim=True)
    trans = torch.rand(B, P, 3, generator=g) - 0.5
    o = torch.randn(B * P, 4, generator=g)
    quat = (o / torch.copysign(o.norm(dim=1), o[:, 0])[:, None]).view(B, P, 4)
    m = valids[..., None]
    batch = {
        'part_pcs': pcs * m[..., None],
        'part_trans': trans * m,
        'part_quat': quat * m,
        'part_valids': valids,
        'data_id': torch.arange(B, dtype=torch.int64),
        'part_ids': torch.arange(P).float().repeat(B, 1) * valids,
    }
    if semantic:
        inst = torch.eye(P).repeat(B, 1, 1) * m
        batch['instance_label'] = inst
        batch['part_label'] = torch.zeros(B, P, 0)
        # groups of geometrically equivalent parts: pairs (1,2), (3,4,5) when valid
        match = torch.zeros(B, P)
        for b, v in enumerate(nv):
            if v >= 3:
                match[b, 1:3] = 1
                pcs_b = batch['part_pcs'][b]
                pcs_b[2] = pcs_b[1]
            if v >= 6:
                match[b, 3:6] = 2
                pcs_b[4] = pcs_b[3]
                pcs_b[5] = pcs_b[3]
        batch['match_ids'] = match
    else:
        batch['instance_label'] = torch.zeros(B, P, 0)
        batch['part_label'] = torch.zeros(B, P, 0)
    if valid_matrix:
        batch['valid_matrix'] = valids[:, :, None] * valids[:, None, :]
    if pin_memory:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device != 'cpu':
        batch = {k: v.to(device, non_blocking=True) for k, v in batch.items()}
    return batch
