"""Dispatch of the model-side hot ops to the sm_100a library.

Every function here is what a reference module's forward used to be, and takes
the module's own parameters, so the nn.Module classes stay plain parameter
containers (checkpoint compatible).

  encode_parts        <- _extract_part_feats (pn_transformer/network.py:59-68)
  pointnet_forward    <- PointNet.forward    (modules/encoder/pointnet.py:29-41)
  dgcnn_forward       <- DGCNN.forward       (modules/encoder/dgcnn.py:77-109)
  transformer_forward <- nn.TransformerEncoder (pn_transformer/transformer.py:63-79)
"""
import torch
import torch.nn.functional as F

from . import _lib


def encode_parts(encoder, part_pcs, part_valids, feat_dim):
    """[B, P, N, 3], [B, P] -> [B, P, C]: run the shared encoder on the valid
    parts only (BatchNorm statistics must exclude padding) and scatter the
    features back; padded parts get zeros.

    The reference indexes with a boolean mask, which synchronises the host to
    learn the output size.  Valid parts are a prefix of the P slots in every
    dataset of the reference (geometry_data.py:101-107), so when the caller
    passes `n_valid` hints we could skip that; in general we keep the mask
    semantics but do the compaction with one nonzero() call reused for gather
    and scatter."""
    B, P, N, _ = part_pcs.shape
    valid_mask = part_valids == 1
    if bool(valid_mask.all()):
        feats = encoder(part_pcs.reshape(B * P, N, 3))
        return feats.view(B, P, -1)
    idx = valid_mask.reshape(-1).nonzero(as_tuple=True)[0]
    valid_feats = encoder(part_pcs.reshape(B * P, N, 3).index_select(0, idx))
    pc_feats = torch.zeros(B * P, feat_dim, dtype=valid_feats.dtype, device=valid_feats.device)
    pc_feats = pc_feats.index_copy(0, idx, valid_feats)
    return pc_feats.view(B, P, feat_dim)


# ---------------------------------------------------------------------------
# PointNet
# ---------------------------------------------------------------------------
def pointnet_forward(x, convs, bns, training, global_feat=True):
    """x [n, N, 3] -> [n, F] (max over points) or [n, N, F]."""
    _lib.require_cuda(x)
    h = x.transpose(2, 1)
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        h = bn(conv(h))
        if i < 4:
            h = F.relu(h)
    return h.max(dim=-1)[0] if global_feat else h.transpose(2, 1).contiguous()


# ---------------------------------------------------------------------------
# DGCNN
# ---------------------------------------------------------------------------
def _graph_feature(x, k):
    n, C, N = x.shape
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x**2, dim=1, keepdim=True)
    idx = (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]
    xt = x.transpose(2, 1)
    nbr = torch.gather(xt.unsqueeze(1).expand(n, N, N, C), 2, idx.unsqueeze(-1).expand(n, N, k, C))
    ctr = xt.unsqueeze(2).expand(n, N, k, C)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()


def dgcnn_forward(x, m, training, k=20):
    """x [n, N, 3] -> [n, F] / [n, N, F]; `m` is the DGCNN module."""
    _lib.require_cuda(x)
    h = x.transpose(2, 1).contiguous()
    feats = []
    for conv in (m.conv1, m.conv2, m.conv3, m.conv4):
        h = conv(_graph_feature(h, k)).max(dim=-1)[0]
        feats.append(h)
    h = m.conv5(torch.cat(feats, dim=1))
    if not m.global_feat:
        return h.transpose(2, 1).contiguous()
    g = torch.cat((h.max(dim=-1)[0], h.mean(dim=-1)), 1)
    return m.out_fc(g)


# ---------------------------------------------------------------------------
# Transformer encoder
# ---------------------------------------------------------------------------
def transformer_forward(tokens, valid_masks, encoder, num_heads, training, dropout):
    _lib.require_cuda(tokens)
    pad = None if valid_masks is None else ~valid_masks
    return encoder(tokens, src_key_padding_mask=pad)
