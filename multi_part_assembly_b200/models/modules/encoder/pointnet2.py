"""PointNet++ part encoders (single- and multi-scale grouping): same parameters, state-dict keys
and forward contract as the reference (models/modules/encoder/pointnet2/pointnet2_ssg.py:6-66,
pointnet2_msg.py:8-43 on top of pointnet2_ops' PointnetSAModule[MSG], pointnet2_modules.py:10-147):
three set-abstraction levels -- furthest point sampling, ball query, grouping, a shared MLP of
1x1 Conv2d (no bias) + BatchNorm2d + ReLU, max over the group -- the last level grouping all
points.  The modules only hold the parameters (`SA_modules.<i>.mlps.<j>.<0|1|3|4|6|7>.*` like the
reference's nn.Sequential); the forward runs the sm_100a kernels of `kernels.pointnet2_forward`."""
import torch.nn as nn

from .... import kernels


def build_shared_mlp(mlp_spec, bn=True):
    """pointnet2_modules.py:10-22."""
    layers = []
    for i in range(1, len(mlp_spec)):
        layers.append(nn.Conv2d(mlp_spec[i - 1], mlp_spec[i], kernel_size=1, bias=not bn))
        if bn:
            layers.append(nn.BatchNorm2d(mlp_spec[i]))
        layers.append(nn.ReLU(True))
    return nn.Sequential(*layers)


class SAModule(nn.Module):
    """Parameter container of one set-abstraction level (PointnetSAModuleMSG,
    pointnet2_modules.py:78-116; a single scale is the one-element case, :119-147).
    npoint None = group all points."""

    def __init__(self, npoint, radii, nsamples, mlps, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps) and use_xyz
        self.npoint, self.radii, self.nsamples = npoint, list(radii), list(nsamples)
        self.groupers = nn.ModuleList()  # parameter-free in the reference; kept for the key layout
        self.mlps = nn.ModuleList()
        for spec in mlps:
            spec = list(spec)
            spec[0] += 3  # use_xyz
            self.mlps.append(build_shared_mlp(spec))


class PointNet2SSG(nn.Module):
    """Input [n, N, 3]; output [n, feat_dim]."""

    def __init__(self, feat_dim):
        super().__init__()
        self.feat_dim = feat_dim
        self._build_model()

    def _build_model(self):
        self.SA_modules = nn.ModuleList([
            SAModule(512, [0.2], [64], [[0, 64, 64, 128]]),
            SAModule(128, [0.4], [64], [[128, 128, 128, 256]]),
            SAModule(None, [None], [None], [[256, 256, 512, self.feat_dim]]),
        ])

    def forward(self, pointcloud):
        return kernels.pointnet2_forward(pointcloud, self, self.training)


class PointNet2MSG(PointNet2SSG):

    def _build_model(self):
        c1 = 64 + 128 + 128
        self.SA_modules = nn.ModuleList([
            SAModule(512, [0.1, 0.2, 0.4], [16, 32, 128],
                     [[0, 32, 32, 64], [0, 64, 64, 128], [0, 64, 96, 128]]),
            SAModule(128, [0.2, 0.4, 0.8], [32, 64, 128],
                     [[c1, 64, 64, 128], [c1, 128, 128, 256], [c1, 128, 128, 256]]),
            SAModule(None, [None], [None], [[128 + 256 + 256, 256, 512, self.feat_dim]]),
        ])
