"""Host-side API surface (no GPU): the reference's config files load
unmodified through `compat`, the built-in configs equal them, and the model
registry builds every hot-path model with the reference's parameter names."""
import glob
import importlib.util
import os

import pytest
import torch

import multi_part_assembly_b200.compat as compat
from multi_part_assembly_b200.configs import get_cfg
from multi_part_assembly_b200.models import build_model

compat.install()
REF_CFG = '/root/reference/configs'
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_CFG), reason='reference tree not present')


def _load(path):
    spec = importlib.util.spec_from_file_location('ref_cfg_' + str(abs(hash(path))), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_cfg_defaults()


def _plain(node):
    return {k: _plain(v) if hasattr(v, 'items') else (list(v) if isinstance(v, (tuple, list)) else v)
            for k, v in node.items()}


CASES = [
    ('pn_transformer/pn_transformer/pn_transformer-32x1-cosine_400e-everyday.py', 'pn_transformer', 'everyday'),
    ('pn_transformer/pn_transformer/pn_transformer-32x1-cosine_400e-partnet_chair.py', 'pn_transformer', 'partnet_chair'),
    ('pn_transformer/pn_transformer_refine/pn_transformer_refine-32x1-cosine_400e-everyday.py', 'pn_transformer_refine', 'everyday'),
    ('dgl/dgl-32x1-cosine_200e-everyday.py', 'dgl', 'everyday'),
    ('dgl/dgl-32x1-cosine_300e-partnet_chair.py', 'dgl', 'partnet_chair'),
    ('global/global-32x1-cosine_200e-everyday.py', 'global', 'everyday'),
    ('global/global-32x1-cosine_200e-partnet_chair.py', 'global', 'partnet_chair'),
]


@needs_ref
@pytest.mark.parametrize('path,model,dataset', CASES)
def test_builtin_cfg_equals_reference_file(path, model, dataset):
    ref = _plain(_load(os.path.join(REF_CFG, path)))
    ours = _plain(get_cfg(model, dataset))
    assert ours == ref


@needs_ref
def test_every_hot_path_reference_config_builds_a_model():
    n = 0
    for sub in ('pn_transformer', 'dgl', 'global'):
        for path in sorted(glob.glob(os.path.join(REF_CFG, sub, '**', '*.py'), recursive=True)):
            cfg = _load(path)
            cfg.freeze()
            model = build_model(cfg)
            assert sum(p.numel() for p in model.parameters()) > 0
            n += 1
    assert n >= 9


def test_parameter_names_and_counts():
    """SURVEY.md appendix A (probe of the reference): parameter counts and the
    state_dict keys a reference checkpoint carries."""
    m = build_model(get_cfg('pn_transformer'))
    assert sum(p.numel() for p in m.parameters()) == 3309639
    sd = m.state_dict()
    for k, shape in {
            'encoder.conv1.weight': (64, 3, 1), 'encoder.conv5.weight': (256, 128, 1),
            'encoder.bn5.running_var': (256, ),
            'corr_module.transformer_encoder.layers.3.self_attn.in_proj_weight': (768, 256),
            'corr_module.transformer_encoder.layers.0.linear1.weight': (1024, 256),
            'corr_module.transformer_encoder.norm.weight': (256, ),
            'pose_predictor.fc_layers.2.weight': (128, 256),
            'pose_predictor.rot_head.weight': (4, 128), 'pose_predictor.trans_head.bias': (3, )}.items():
        assert tuple(sd[k].shape) == shape, k
    assert sum(p.numel() for p in build_model(get_cfg('dgl')).parameters()) == 3245782
    assert sum(p.numel() for p in build_model(get_cfg('global')).parameters()) == 167303
    assert sum(p.numel() for p in build_model(get_cfg('pn_transformer_refine')).parameters()) == 1595477
    cfg = get_cfg('dgl', encoder='dgcnn')
    sd = build_model(cfg).state_dict()
    # DGCNN registers each BatchNorm twice (dgcnn.py:51-59)
    assert 'encoder.bn1.weight' in sd and 'encoder.conv1.1.weight' in sd
    assert tuple(sd['encoder.conv4.0.weight'].shape) == (256, 256, 1, 1)
    assert tuple(sd['encoder.out_fc.weight'].shape) == (128, 256)


def test_unsupported_models_raise():
    cfg = get_cfg('pn_transformer')
    cfg.model.name = 'lstm'
    with pytest.raises(NotImplementedError):
        build_model(cfg)
    cfg.model.name = 'pn_transformer'
    cfg.model.encoder = 'pointnet3'
    with pytest.raises(NotImplementedError):
        build_model(cfg)


def test_pointnet2_encoders_build():
    """Encoder registry values 'pointnet2_ssg' / 'pointnet2_msg' (reference encoder/__init__.py:
    11-19) with the reference's state-dict layout (nn.Sequential of conv / bn / relu per scale)."""
    from multi_part_assembly_b200.models import build_encoder
    ssg = build_encoder('pointnet2_ssg', 256)
    sd = ssg.state_dict()
    assert tuple(sd['SA_modules.0.mlps.0.0.weight'].shape) == (64, 3, 1, 1)
    assert tuple(sd['SA_modules.1.mlps.0.6.weight'].shape) == (256, 128, 1, 1)
    assert tuple(sd['SA_modules.2.mlps.0.6.weight'].shape) == (256, 512, 1, 1)
    assert 'SA_modules.0.mlps.0.1.running_mean' in sd and 'SA_modules.0.mlps.0.0.bias' not in sd
    msg = build_encoder('pointnet2_msg', 128)
    sd = msg.state_dict()
    assert tuple(sd['SA_modules.0.mlps.2.3.weight'].shape) == (96, 64, 1, 1)
    assert tuple(sd['SA_modules.1.mlps.1.0.weight'].shape) == (128, 64 + 128 + 128 + 3, 1, 1)
    assert tuple(sd['SA_modules.2.mlps.0.0.weight'].shape) == (256, 128 + 256 + 256 + 3, 1, 1)
    cfg = get_cfg('pn_transformer')
    cfg.model.encoder = 'pointnet2_ssg'
    assert build_model(cfg) is not None


def test_cfgnode_semantics():
    cfg = get_cfg('pn_transformer')
    assert cfg.loss.get('sample_iter', 1) == 1 and 'noise_dim' in cfg.loss
    c2 = cfg.clone()
    c2.model.encoder = 'dgcnn'
    assert cfg.model.encoder == 'pointnet'
    c2.freeze()
    with pytest.raises(AttributeError):
        c2.model.encoder = 'pointnet'
    c2.defrost()
    c2.merge_from_list(['exp.batch_size', 4])
    assert c2.exp.batch_size == 4 and cfg.optimizer.clip_grad is None


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA tensors."""
    from multi_part_assembly_b200.utils import chamfer_distance, transform_pc, Rotation3D
    with pytest.raises(RuntimeError):
        chamfer_distance(torch.rand(1, 8, 3), torch.rand(1, 8, 3))
    with pytest.raises(RuntimeError):
        transform_pc(torch.rand(2, 3), Rotation3D(torch.rand(2, 4) + 1), torch.rand(2, 5, 3))
