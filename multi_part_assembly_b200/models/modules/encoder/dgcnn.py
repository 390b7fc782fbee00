"""DGCNN part encoder: same parameters (incl. the doubled BatchNorm keys
`bnK` / `convK.1`) and forward contract as the reference
(models/modules/encoder/dgcnn.py:41-109): four EdgeConv layers over a k=20
k-NN graph rebuilt in feature space before each layer, conv5 on the
concatenated 512 channels, max+avg pooling, Linear."""
import torch.nn as nn

from .... import kernels


class DGCNN(nn.Module):
    """Input [n, N, 3]; output [n, feat_dim] (global_feat) or [n, N, feat_dim]."""

    K = 20

    def __init__(self, feat_dim, global_feat=True):
        super().__init__()
        chans = [(6, 64), (128, 64), (128, 128), (256, 256)]
        for i, (_, co) in enumerate(chans):
            setattr(self, f'bn{i + 1}', nn.BatchNorm2d(co))
        self.bn5 = nn.BatchNorm1d(feat_dim)
        for i, (ci, co) in enumerate(chans):
            setattr(self, f'conv{i + 1}', nn.Sequential(
                nn.Conv2d(ci, co, kernel_size=1, bias=False),
                getattr(self, f'bn{i + 1}'), nn.LeakyReLU(negative_slope=0.2)))
        self.conv5 = nn.Sequential(
            nn.Conv1d(512, feat_dim, kernel_size=1, bias=False), self.bn5,
            nn.LeakyReLU(negative_slope=0.2))
        self.global_feat = global_feat
        if global_feat:
            self.out_fc = nn.Linear(feat_dim * 2, feat_dim)

    supports_valids = True

    def forward(self, x, valids=None):
        """x [n, N, 3]; `valids` [n] (extension): parts flagged 0 are masked on the device
        (zero features, no BatchNorm contribution) instead of being compacted on the host."""
        return kernels.dgcnn_forward(x, self, self.training, self.K, valids)
