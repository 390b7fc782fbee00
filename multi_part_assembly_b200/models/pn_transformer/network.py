"""PointNet + Transformer assembly model (reference
models/pn_transformer/network.py:9-139)."""
import torch

from ..modules.base_model import BaseModel
from ..modules.encoder import build_encoder
from ..modules.regressor import StocasticPoseRegressor
from ... import kernels
from .transformer import TransformerEncoder


class PNTransformer(BaseModel):
    """Encoder: shared per-part PointNet; correlator: TransformerEncoder over
    the part tokens; predictor: stochastic MLP pose head."""

    def __init__(self, cfg):
        super().__init__(cfg)
        self.encoder = self._init_encoder()
        self.corr_module = self._init_corr_module()
        self.pose_predictor = self._init_pose_predictor()

    def _init_encoder(self):
        return build_encoder(self.cfg.model.encoder, feat_dim=self.pc_feat_dim, global_feat=True)

    def _init_corr_module(self):
        m = self.cfg.model
        return TransformerEncoder(
            d_model=self.pc_feat_dim, num_heads=m.transformer_heads,
            ffn_dim=m.transformer_feat_dim, num_layers=m.transformer_layers,
            norm_first=m.transformer_pre_ln)

    def _pose_in_dim(self):
        dim = self.pc_feat_dim
        if self.semantic:
            dim += self.max_num_part
        if self.use_part_label:
            dim += self.cfg.data.num_part_category
        return dim

    def _init_pose_predictor(self):
        return StocasticPoseRegressor(feat_dim=self._pose_in_dim(),
                                      noise_dim=self.cfg.loss.noise_dim, rot_type=self.rot_type)

    def _extract_part_feats(self, part_pcs, part_valids, valid_mask=None):
        """[B, P, N, 3] -> [B, P, C]; padded parts get zero features and are
        excluded from the encoder's BatchNorm statistics (reference :59-68)."""
        return kernels.encode_parts(self.encoder, part_pcs, part_valids, self.pc_feat_dim,
                                    valid_mask=valid_mask)

    def forward(self, data_dict):
        """data_dict: part_pcs [B,P,N,3], part_valids [B,P], part_label
        [B,P,L], instance_label [B,P,I]; optionally the cached
        `pre_pose_feats`.  Returns rot (Rotation3D), trans, pre_pose_feats."""
        feats = data_dict.get('pre_pose_feats', None)
        if feats is None:
            part_valids = data_dict['part_valids']
            valid_mask = part_valids == 1  # computed once for the encoder and the transformer
            pc_feats = self._extract_part_feats(data_dict['part_pcs'], part_valids, valid_mask)
            corr_feats = self.corr_module(pc_feats, valid_mask)
            feats = torch.cat([corr_feats, data_dict['part_label'].type_as(corr_feats),
                               data_dict['instance_label'].type_as(corr_feats)], dim=-1)
        rot, trans = self.pose_predictor(feats)
        return {'rot': self._wrap_rotation(rot), 'trans': trans, 'pre_pose_feats': feats}

    def _loss_function(self, data_dict, out_dict={}, optimizer_idx=-1):
        """One Min-of-N sample: predict, then `_calc_loss`; the features before
        the stochastic pose head are cached across samples (reference :106-139)."""
        forward_dict = {k: data_dict[k] for k in
                        ('part_pcs', 'part_valids', 'part_label', 'instance_label')}
        forward_dict['pre_pose_feats'] = out_dict.get('pre_pose_feats', None)
        pred = self.forward(forward_dict)
        loss_dict, out_dict = self._calc_loss(pred, data_dict)
        out_dict['pre_pose_feats'] = pred['pre_pose_feats']
        return loss_dict, out_dict
