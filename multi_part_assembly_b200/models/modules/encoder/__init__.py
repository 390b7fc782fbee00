from .pointnet import PointNet
from .dgcnn import DGCNN


def build_encoder(arch, feat_dim, global_feat=True, **kwargs):
    """Encoder registry (reference models/modules/encoder/__init__.py:6-21)."""
    if arch == 'pointnet':
        return PointNet(feat_dim, global_feat=global_feat)
    if arch == 'dgcnn':
        return DGCNN(feat_dim, global_feat=global_feat)
    if 'pointnet2' in arch:
        raise NotImplementedError(
            f'{arch}: the PointNet++ encoders are outside the B200 hot-path scope '
            '(SURVEY.md 8f rank 3); no shipped config selects them')
    raise NotImplementedError(f'{arch} is not supported')
