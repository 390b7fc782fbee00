from .pointnet import PointNet
from .dgcnn import DGCNN
from .pointnet2 import PointNet2SSG, PointNet2MSG


def build_encoder(arch, feat_dim, global_feat=True, **kwargs):
    """Encoder registry (reference models/modules/encoder/__init__.py:6-21)."""
    if arch == 'pointnet':
        return PointNet(feat_dim, global_feat=global_feat)
    if arch == 'dgcnn':
        return DGCNN(feat_dim, global_feat=global_feat)
    if 'pointnet2' in arch:
        assert global_feat
        if 'ssg' in arch:
            return PointNet2SSG(feat_dim)
        if 'msg' in arch:
            return PointNet2MSG(feat_dim)
        raise NotImplementedError(f'{arch} not supported')
    raise NotImplementedError(f'{arch} is not supported')
