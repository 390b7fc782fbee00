"""Plain-PyTorch (CPU, fp32/fp64) restatement of the floating-point modules on
the hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Everything is written functionally over a `state_dict`-style mapping of
parameter tensors so the same weights can be fed to the CUDA product modules
and to this oracle.  Reference paths are relative to
/root/reference/multi_part_assembly.  Pinned by tests/golden/*.npz, which
oracle/make_golden.py generates from the reference modules themselves.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cpu as _cpu


# --------------------------------------------------------------------------
# pytorch3d semantics (third-party, absent; PARITY UNPINNED by reference tests)
# --------------------------------------------------------------------------
def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack((ow, ox, oy, oz), -1)


def quaternion_invert(q):
    return q * torch.tensor([1, -1, -1, -1], dtype=q.dtype, device=q.device)


def quaternion_apply(q, point):
    real = point.new_zeros(point.shape[:-1] + (1, ))
    out = quaternion_raw_multiply(
        quaternion_raw_multiply(q, torch.cat((real, point), -1)), quaternion_invert(q))
    return out[..., 1:]


def qrot(q, v):
    """utils/transforms.py:75-87."""
    if q.dim() == v.dim() - 1:
        q = q.unsqueeze(-2).repeat_interleave(v.shape[-2], dim=-2)
    return quaternion_apply(q, v)


def qtransform(t, q, v):
    """utils/transforms.py:90-109."""
    if t.dim() == v.dim() - 1:
        t = t.unsqueeze(-2).repeat_interleave(v.shape[-2], dim=-2)
    return qrot(q, v) + t


def process_zero_quat(q):
    """utils/rotation.py:121-128."""
    q = q.float()
    keep = torch.norm(q, p=2, dim=-1, keepdim=True).abs() > 0.5
    iden = torch.zeros_like(q)
    iden[..., 0] = 1.
    return torch.where(keep, q, iden)


# --------------------------------------------------------------------------
# Chamfer (C oracle underneath) with autograd, CPU
# --------------------------------------------------------------------------
class _Chamfer(torch.autograd.Function):

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, i1, d2, i2 = _cpu.chamfer_forward(xyz1.detach().float().numpy(),
                                              xyz2.detach().float().numpy())
        i1, i2 = torch.from_numpy(i1), torch.from_numpy(i2)
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        return torch.from_numpy(d1), torch.from_numpy(d2)

    @staticmethod
    def backward(ctx, g1, g2):
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        a, b = _cpu.chamfer_backward(g1.float().numpy(), g2.float().numpy(),
                                     xyz1.detach().float().numpy(),
                                     xyz2.detach().float().numpy(), i1.numpy(), i2.numpy())
        return torch.from_numpy(a), torch.from_numpy(b)


def chamfer_distance(xyz1, xyz2):
    """utils/chamfer/chamfer.py:36-64 (BNC layout, squared distances)."""
    return _Chamfer.apply(xyz1.contiguous(), xyz2.contiguous())


# --------------------------------------------------------------------------
# losses (utils/loss.py)
# --------------------------------------------------------------------------
def valid_mean(x, valids):
    valids = valids.float()
    return (x * valids).sum(1) / valids.sum(1)


def trans_l2_loss(t1, t2, valids):
    return valid_mean((t1 - t2).pow(2).sum(-1), valids)


def rot_cosine_loss(q1, q2, valids):
    return valid_mean(1. - torch.abs(torch.sum(q1 * q2, dim=-1)), valids)


def rot_points_l2_loss(pts, q1, q2, valids):
    return valid_mean((qrot(q1, pts) - qrot(q2, pts)).pow(2).sum(-1).mean(-1), valids)


def rot_points_cd_loss(pts, q1, q2, valids):
    B = pts.shape[0]
    d1, d2 = chamfer_distance(qrot(q1, pts).flatten(0, 1), qrot(q2, pts).flatten(0, 1))
    return valid_mean((d1.mean(1) + d2.mean(1)).view(B, -1), valids)


def shape_cd_loss(pts, t1, t2, q1, q2, valids, training=True):
    B, P, N, _ = pts.shape
    pts = pts.detach().clone().masked_fill(valids[..., None, None] == 0, 1e3)
    d1, d2 = chamfer_distance(qtransform(t1, q1, pts).flatten(1, 2),
                              qtransform(t2, q2, pts).flatten(1, 2))
    valids = valids.float()
    if training:
        rep = valids.unsqueeze(2).repeat(1, 1, N).view(B, -1)
        return (d1 * rep).mean(1) + (d2 * rep).mean(1)
    return valid_mean((d1 + d2).view(B, P, N).mean(-1), valids)


def calc_part_acc(pts, t1, t2, q1, q2, valids):
    """utils/eval_utils.py:13-46."""
    B, P = pts.shape[:2]
    d1, d2 = chamfer_distance(qtransform(t1, q1, pts).flatten(0, 1),
                              qtransform(t2, q2, pts).flatten(0, 1))
    cd = (d1.mean(1) + d2.mean(1)).view(B, P)
    acc = (cd < 0.01) & (valids == 1)
    return acc.sum(-1) / (valids == 1).sum(-1)


# --------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------
def _bn(x, sd, name, training, eps=1e-5):
    return F.batch_norm(x, None if training else sd[f'{name}.running_mean'],
                        None if training else sd[f'{name}.running_var'],
                        sd[f'{name}.weight'], sd[f'{name}.bias'], training=training,
                        momentum=0.0, eps=eps)


def pointnet_forward(x, sd, training=True, global_feat=True, prefix=''):
    """models/modules/encoder/pointnet.py:29-41.  x [n, N, 3] -> [n, F]."""
    h = x.transpose(2, 1)
    for i in range(1, 6):
        h = F.conv1d(h, sd[f'{prefix}conv{i}.weight'])
        h = _bn(h, sd, f'{prefix}bn{i}', training)
        if i < 5:
            h = F.relu(h)
    return h.max(dim=-1)[0] if global_feat else h.transpose(2, 1).contiguous()


def knn_scores(x):
    """models/modules/encoder/dgcnn.py:10-12; x [n, C, N] -> [n, N, N]."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x**2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def knn(x, k):
    return knn_scores(x).topk(k=k, dim=-1)[1]


def graph_feature(x, k=20, idx=None):
    """models/modules/encoder/dgcnn.py:18-38; x [n, C, N] -> [n, 2C, N, k]."""
    n, C, N = x.shape
    if idx is None:
        idx = knn(x, k)
    xt = x.transpose(2, 1)  # [n, N, C]
    nbr = torch.gather(xt.unsqueeze(1).expand(n, N, N, C), 2,
                       idx.unsqueeze(-1).expand(n, N, k, C))  # [n, N, k, C]
    ctr = xt.unsqueeze(2).expand(n, N, k, C)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()


def dgcnn_forward(x, sd, training=True, k=20, prefix='', return_idx=False):
    """models/modules/encoder/dgcnn.py:77-109.  x [n, N, 3] -> [n, F]."""
    h = x.transpose(2, 1).contiguous()
    feats, idxs = [], []
    for i in range(1, 5):
        idx = knn(h, k)
        idxs.append(idx)
        e = graph_feature(h, k, idx)
        e = F.conv2d(e, sd[f'{prefix}conv{i}.0.weight'])
        e = F.leaky_relu(_bn(e, sd, f'{prefix}bn{i}', training), 0.2)
        h = e.max(dim=-1)[0]
        feats.append(h)
    h = torch.cat(feats, dim=1)
    h = F.conv1d(h, sd[f'{prefix}conv5.0.weight'])
    h = F.leaky_relu(_bn(h, sd, f'{prefix}bn5', training), 0.2)
    g = torch.cat((h.max(dim=-1)[0], h.mean(dim=-1)), 1)
    out = F.linear(g, sd[f'{prefix}out_fc.weight'], sd[f'{prefix}out_fc.bias'])
    return (out, idxs) if return_idx else out


# --------------------------------------------------------------------------
# transformer encoder (nn.TransformerEncoder, pre-LN, batch_first, ReLU, eval /
# dropout 0), models/pn_transformer/transformer.py:4-79
# --------------------------------------------------------------------------
def transformer_forward(tokens, valid_mask, sd, num_heads, num_layers, prefix='',
                        norm_first=True):
    B, S, D = tokens.shape
    hd = D // num_heads
    x = tokens
    neg = None
    if valid_mask is not None:
        neg = torch.zeros(B, 1, 1, S, dtype=x.dtype)
        neg.masked_fill_(~valid_mask.view(B, 1, 1, S), float('-inf'))

    def attn(h, p):
        qkv = F.linear(h, sd[p + 'self_attn.in_proj_weight'], sd[p + 'self_attn.in_proj_bias'])
        q, k, v = qkv.view(B, S, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
        if neg is not None:
            s = s + neg
        o = torch.matmul(torch.softmax(s, dim=-1), v)  # [B, H, S, hd]
        o = o.permute(0, 2, 1, 3).reshape(B, S, D)
        return F.linear(o, sd[p + 'self_attn.out_proj.weight'], sd[p + 'self_attn.out_proj.bias'])

    def ffn(h, p):
        return F.linear(F.relu(F.linear(h, sd[p + 'linear1.weight'], sd[p + 'linear1.bias'])),
                        sd[p + 'linear2.weight'], sd[p + 'linear2.bias'])

    def ln(h, p):
        return F.layer_norm(h, (D, ), sd[p + '.weight'], sd[p + '.bias'], 1e-5)

    for l in range(num_layers):
        p = f'{prefix}transformer_encoder.layers.{l}.'
        if norm_first:
            x = x + attn(ln(x, p + 'norm1'), p)
            x = x + ffn(ln(x, p + 'norm2'), p)
        else:
            x = ln(x + attn(x, p), p + 'norm1')
            x = ln(x + ffn(x, p), p + 'norm2')
    if norm_first:
        x = ln(x, f'{prefix}transformer_encoder.norm')
    if f'{prefix}out_fc.weight' in sd:
        x = F.linear(x, sd[f'{prefix}out_fc.weight'], sd[f'{prefix}out_fc.bias'])
    return x


def pose_regressor_forward(x, sd, prefix='', noise=None):
    """models/modules/regressor.py:58-84 (quat head)."""
    if noise is not None:
        x = torch.cat([x, noise], dim=-1)
    f = F.leaky_relu(F.linear(x, sd[prefix + 'fc_layers.0.weight'], sd[prefix + 'fc_layers.0.bias']), 0.2)
    f = F.leaky_relu(F.linear(f, sd[prefix + 'fc_layers.2.weight'], sd[prefix + 'fc_layers.2.bias']), 0.2)
    rot = F.normalize(F.linear(f, sd[prefix + 'rot_head.weight'], sd[prefix + 'rot_head.bias']), p=2, dim=-1)
    trans = F.linear(f, sd[prefix + 'trans_head.weight'], sd[prefix + 'trans_head.bias'])
    return rot, trans


# --------------------------------------------------------------------------
# pn_transformer forward + loss (geometric config), the bench "step"
# models/pn_transformer/network.py:59-139 + models/modules/base_model.py:240-387
# --------------------------------------------------------------------------
GEOMETRIC_LOSS_W = dict(trans_loss=1., rot_pt_cd_loss=10., transform_pt_cd_loss=10.,
                        rot_loss=0.2, rot_pt_l2_loss=1.)


def pn_transformer_forward(batch, sd, num_heads=8, num_layers=4, training=True, noise=None):
    pcs, valids = batch['part_pcs'], batch['part_valids']
    B, P, N, _ = pcs.shape
    mask = valids == 1
    feats_valid = pointnet_forward(pcs[mask], sd, training=training, prefix='encoder.')
    C = feats_valid.shape[-1]
    feats = torch.zeros(B, P, C, dtype=feats_valid.dtype)
    feats[mask] = feats_valid
    corr = transformer_forward(feats, mask, sd, num_heads, num_layers, prefix='corr_module.')
    x = torch.cat([corr, batch['part_label'].type_as(corr), batch['instance_label'].type_as(corr)], -1)
    rot, trans = pose_regressor_forward(x, sd, prefix='pose_predictor.', noise=noise)
    return process_zero_quat(rot), trans


def geometric_losses(batch, pred_rot, pred_trans, training=True, weights=GEOMETRIC_LOSS_W):
    """base_model.py:240-314 (geometric branch) + weighted total (:367-373);
    returns per-term [B] tensors and the scalar mean loss."""
    pcs, valids = batch['part_pcs'], batch['part_valids']
    gt_t = batch['part_trans']
    gt_q = process_zero_quat(batch['part_quat'])
    terms = {
        'trans_loss': trans_l2_loss(pred_trans, gt_t, valids),
        'rot_pt_cd_loss': rot_points_cd_loss(pcs, pred_rot, gt_q, valids),
        'transform_pt_cd_loss': shape_cd_loss(pcs, pred_trans, gt_t, pred_rot, gt_q, valids,
                                              training=training),
        'rot_loss': rot_cosine_loss(pred_rot, gt_q, valids),
        'rot_pt_l2_loss': rot_points_l2_loss(pcs, pred_rot, gt_q, valids),
    }
    total = sum(terms[k] * weights[k] for k in terms)
    out = {k: v.mean() for k, v in terms.items()}
    out['loss'] = total.mean()
    return out, terms


def numpy_batch(batch):
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
