"""Pose heads (reference models/modules/regressor.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def normalize_rot6d(rot):
    """Gram-Schmidt the two 3-vectors of a 6D rotation ([..., 6] or [..., 2, 3])."""
    unflatten = rot.shape[-1] == 3
    if unflatten:
        rot = rot.flatten(-2, -1)
    a1, a2 = rot[..., :3], rot[..., 3:]
    b1 = F.normalize(a1, p=2, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, p=2, dim=-1)
    rot = torch.cat([b1, b2], dim=-1)
    return rot.unflatten(-1, (2, 3)) if unflatten else rot


class PoseRegressor(nn.Module):
    """feat -> 256 -> 128 -> {rotation head (L2-normalised quat or 6D), translation head}."""

    def __init__(self, feat_dim, rot_type='quat', norm_rot=True):
        super().__init__()
        if rot_type not in ('quat', 'rmat'):
            raise NotImplementedError(f'rotation {rot_type} is not supported')
        self.rot_type = rot_type
        self.norm_rot = norm_rot
        self.fc_layers = nn.Sequential(
            nn.Linear(feat_dim, 256), nn.LeakyReLU(0.2),
            nn.Linear(256, 128), nn.LeakyReLU(0.2))
        self.rot_head = nn.Linear(128, 4 if rot_type == 'quat' else 6)
        self.trans_head = nn.Linear(128, 3)

    def forward(self, x):
        from ... import kernels
        if self.rot_type == 'quat' and x.is_cuda and not torch.is_grad_enabled():
            return kernels.pose_head_forward(x, self)  # native (fp32), forward only
        f = self.fc_layers(x)
        rot = self.rot_head(f)
        if self.norm_rot:
            rot = F.normalize(rot, p=2, dim=-1) if self.rot_type == 'quat' \
                else normalize_rot6d(rot)
        return rot, self.trans_head(f)


class StocasticPoseRegressor(PoseRegressor):
    """PoseRegressor with `noise_dim` Gaussian inputs appended (MoN sampling)."""

    def __init__(self, feat_dim, noise_dim, rot_type='quat', norm_rot=True):
        super().__init__(feat_dim + noise_dim, rot_type, norm_rot)
        self.noise_dim = noise_dim

    def forward(self, x):
        if self.noise_dim == 0:  # deterministic head (geometric configs): nothing to sample
            return super().forward(x)
        # CPU generator, like the reference (regressor.py:82), so seeds reproduce its samples
        noise = torch.randn(list(x.shape[:-1]) + [self.noise_dim]).type_as(x)
        return super().forward(torch.cat([x, noise], dim=-1))
