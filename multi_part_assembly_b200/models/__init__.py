"""Model registry (reference models/__init__.py:10-26)."""
from .modules import *  # noqa: F401,F403
from .modules import BaseModel, build_encoder, PoseRegressor, StocasticPoseRegressor
from .pn_transformer import PNTransformer, PNTransformerRefine
from .b_global import GlobalModel
from .dgl import DGLModel

_MODELS = {
    'global': GlobalModel,
    'dgl': DGLModel,
    'pn_transformer': PNTransformer,
    'pn_transformer_refine': PNTransformerRefine,
}


def build_model(cfg):
    name = cfg.model.name
    if name in _MODELS:
        return _MODELS[name](cfg)
    if name in ('identity', 'lstm', 'rgl_net'):
        raise NotImplementedError(
            f'Model {name}: outside the B200 hot-path scope (SURVEY.md 2.1 #13/#14); '
            'supported: ' + ', '.join(sorted(_MODELS)))
    raise NotImplementedError(f'Model {name} not supported')
