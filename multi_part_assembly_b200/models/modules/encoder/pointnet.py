"""PointNet part encoder: same parameters and forward contract as the
reference (models/modules/encoder/pointnet.py:6-41): five bias-free 1x1 convs
3->64->64->64->128->feat_dim, BatchNorm1d after each, ReLU after the first
four, max over the N points."""
import torch
import torch.nn as nn

from .... import kernels


class PointNet(nn.Module):
    """Input [n, N, 3]; output [n, feat_dim] (global_feat) or [n, N, feat_dim]."""

    def __init__(self, feat_dim, global_feat=True):
        super().__init__()
        dims = [3, 64, 64, 64, 128, feat_dim]
        for i in range(5):
            setattr(self, f'conv{i + 1}', nn.Conv1d(dims[i], dims[i + 1], kernel_size=1, bias=False))
        for i in range(5):
            setattr(self, f'bn{i + 1}', nn.BatchNorm1d(dims[i + 1]))
        self.global_feat = global_feat

    supports_valids = True

    def forward(self, x, valids=None):
        """x [n, N, 3]; `valids` [n] (extension): parts flagged 0 are skipped on
        the device (zero features, no BatchNorm contribution)."""
        convs = [getattr(self, f'conv{i}') for i in range(1, 6)]
        bns = [getattr(self, f'bn{i}') for i in range(1, 6)]
        return kernels.pointnet_forward(x, convs, bns, self.training, self.global_feat, valids)
