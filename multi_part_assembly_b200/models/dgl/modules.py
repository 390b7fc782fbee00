"""MLPs of the dynamic-graph-learning model (reference models/dgl/modules.py).  The modules
hold the reference's parameters; without autograd (evaluation, benchmarking, CUDA-graph replay)
their forward runs on the tensor-core GEMM + BatchNorm kernels of `kernels`, with autograd on
the stock layers."""
import torch
import torch.nn as nn

from ... import kernels


def _native(x):
    return x.is_cuda and not torch.is_grad_enabled()


class _PointwiseMLP3(nn.Module):
    """[M, P, 2F] -> [M, P, F]: three 1x1 conv + BatchNorm1d + ReLU over the
    second axis (BatchNorm statistics include padded rows, as in the reference)."""

    def __init__(self, feat_len):
        super().__init__()
        self.conv1 = nn.Conv1d(2 * feat_len, 512, 1)
        self.conv2 = nn.Conv1d(512, 512, 1)
        self.conv3 = nn.Conv1d(512, feat_len, 1)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(512)
        self.bn3 = nn.BatchNorm1d(feat_len)

    def forward(self, x):
        if _native(x):
            return kernels.conv_bn_relu_rows(
                x, ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)), self.training)
        x = x.permute(0, 2, 1)
        x = torch.relu(self.bn1(self.conv1(x)))
        x = torch.relu(self.bn2(self.conv2(x)))
        x = torch.relu(self.bn3(self.conv3(x)))
        return x.permute(0, 2, 1)


class MLP3(_PointwiseMLP3):
    """Edge MLP."""


class MLP4(_PointwiseMLP3):
    """Node MLP."""


class RelationNet(nn.Module):
    """Pairwise pose features [.., 256] -> relation weight in (0, 1)."""

    def __init__(self):
        super().__init__()
        self.mlp1 = nn.Linear(128 + 128, 256)
        self.mlp2 = nn.Linear(256, 512)
        self.mlp3 = nn.Linear(512, 1)

    def forward(self, x):
        if _native(x):
            return kernels.linear_chain(x, ((self.mlp1, kernels.ACT_RELU), (self.mlp2, kernels.ACT_RELU),
                                            (self.mlp3, kernels.ACT_SIGMOID)))
        x = torch.relu(self.mlp1(x))
        x = torch.relu(self.mlp2(x))
        return torch.sigmoid(self.mlp3(x))


class PoseEncoder(nn.Module):
    """Pose [.., pose_dim] -> feature [.., 128]."""

    def __init__(self, pose_dim):
        super().__init__()
        self.mlp1 = nn.Linear(pose_dim, 256)
        self.mlp2 = nn.Linear(256, 128)

    def forward(self, x):
        if _native(x):
            return kernels.linear_chain(x, ((self.mlp1, kernels.ACT_RELU), (self.mlp2, kernels.ACT_RELU)))
        return torch.relu(self.mlp2(torch.relu(self.mlp1(x))))
