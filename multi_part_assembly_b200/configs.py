"""Built-in equivalents of the reference's shipped configs for the hot-path
models (the reference's own files under configs/ also load unmodified through
`compat`; these exist because /root/reference does not travel to the GPU box).
tests/test_configs.py checks them field by field against the reference files.

    cfg = get_cfg('pn_transformer', dataset='everyday')
"""
from .compat.yacs_config import CfgNode as CN


def _exp(num_epochs=200, val_every=10):
    return CN(dict(ckp_dir='checkpoint/', weight_file='', gpus=[0], num_workers=8,
                   batch_size=32, num_epochs=num_epochs, val_every=val_every, val_sample_vis=5))


def _optimizer(warmup_ratio=0.):
    return CN(dict(lr=1e-3, weight_decay=0., warmup_ratio=warmup_ratio, clip_grad=None,
                   lr_scheduler='cosine', lr_decay_factor=100.))


_COLORS = [[0, 204, 0], [204, 0, 0], [0, 204, 0], [127, 127, 0], [127, 0, 127], [0, 127, 127],
           [76, 153, 0], [153, 0, 76], [76, 0, 153], [153, 76, 0], [76, 0, 153], [153, 0, 76],
           [204, 51, 127], [204, 51, 127], [51, 204, 127], [51, 127, 204], [127, 51, 204],
           [127, 204, 51], [76, 76, 178], [76, 178, 76], [178, 76, 76]]


def _data_everyday(data_keys=('part_ids', )):
    return CN(dict(
        dataset='geometry', data_dir='./data/breaking_bad', data_fn='everyday.{}.txt',
        data_keys=tuple(data_keys), category='', rot_range=-1., num_pc_points=1000,
        min_num_part=2, max_num_part=20, shuffle_parts=False, overfit=-1,
        all_category=['BeerBottle', 'Bowl', 'Cup', 'DrinkingUtensil', 'Mug', 'Plate', 'Spoon',
                      'Teacup', 'ToyFigure', 'WineBottle', 'Bottle', 'Cookie', 'DrinkBottle',
                      'Mirror', 'PillBottle', 'Ring', 'Statue', 'Teapot', 'Vase', 'WineGlass'],
        colors=_COLORS))


def _data_partnet_chair(data_keys=('part_ids', 'match_ids', 'contact_points')):
    return CN(dict(
        dataset='partnet', data_dir='./data/partnet', data_fn='Chair.{}.npy',
        data_keys=tuple(data_keys), category='Chair', num_pc_points=1000, num_part_category=57,
        min_num_part=2,
        max_num_part=20, shuffle_parts=False, overfit=-1, colors=_COLORS))


def _loss_geometric():
    return CN(dict(noise_dim=0, trans_loss_w=1., rot_pt_cd_loss_w=10., transform_pt_cd_loss_w=10.,
                   use_rot_loss=True, rot_loss_w=0.2, use_rot_pt_l2_loss=True,
                   rot_pt_l2_loss_w=1.))


def _loss_semantic():
    return CN(dict(noise_dim=32, sample_iter=5, trans_loss_w=1., rot_pt_cd_loss_w=10.,
                   transform_pt_cd_loss_w=10., use_rot_loss=False, use_rot_pt_l2_loss=False))


def _model(name, encoder='pointnet'):
    if name == 'pn_transformer':
        return CN(dict(name=name, rot_type='quat', pc_feat_dim=256, encoder=encoder,
                       transformer_feat_dim=1024, transformer_heads=8, transformer_layers=4,
                       transformer_pre_ln=True))
    if name == 'pn_transformer_refine':
        return CN(dict(name=name, rot_type='quat', pc_feat_dim=128, encoder=encoder,
                       transformer_pos_enc=(128, 128), transformer_feat_dim=512,
                       transformer_heads=8, transformer_layers=2, transformer_pre_ln=True,
                       pose_pc_feat=True, refine_steps=3))
    if name == 'dgl':
        return CN(dict(name=name, rot_type='quat', pc_feat_dim=128, encoder=encoder, gnn_iter=3,
                       merge_node=True))
    if name == 'global':
        return CN(dict(name=name, rot_type='quat', pc_feat_dim=128, encoder=encoder))
    raise KeyError(name)


def get_cfg(model='pn_transformer', dataset='everyday', encoder='pointnet'):
    """CfgNode with the five sub-nodes exp/data/optimizer/model/loss of
    configs/<model>/<model>-32x1-cosine_*e-<dataset>.py."""
    semantic = dataset == 'partnet_chair'
    cfg = CN()
    cfg.model = _model(model, encoder)
    if model in ('pn_transformer', 'pn_transformer_refine'):
        cfg.exp = _exp(num_epochs=400)
        cfg.optimizer = _optimizer(warmup_ratio=0.05)
        keys = None
    elif model == 'dgl':
        cfg.exp = _exp(num_epochs=300 if semantic else 200, val_every=5)
        cfg.optimizer = _optimizer()
        keys = ('part_ids', 'match_ids', 'contact_points', 'valid_matrix') if semantic \
            else ('part_ids', 'valid_matrix')
        if not semantic:
            cfg.model.merge_node = False
    else:
        cfg.exp = _exp(num_epochs=200)
        cfg.optimizer = _optimizer()
        keys = None
    if semantic:
        cfg.data = _data_partnet_chair(keys) if keys else _data_partnet_chair()
        cfg.loss = _loss_semantic()
    else:
        cfg.data = _data_everyday(keys) if keys else _data_everyday()
        cfg.loss = _loss_geometric()
    return cfg
