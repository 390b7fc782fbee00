"""`B-Global` baseline (reference models/b_global/network.py:7-132): a part
PointNet plus a second PointNet over the whole [B, P*N, 3] cloud (padding
included, as in the reference :56-60), MLP pose head."""
import torch

from ..modules.base_model import BaseModel
from ..modules.encoder import build_encoder
from ..modules.regressor import StocasticPoseRegressor
from ... import kernels


class GlobalModel(BaseModel):

    def __init__(self, cfg):
        super().__init__(cfg)
        self.encoder = self._init_encoder()
        self.global_encoder = self._init_encoder()
        self.pose_predictor = self._init_pose_predictor()

    def _init_encoder(self):
        return build_encoder(self.cfg.model.encoder, feat_dim=self.pc_feat_dim, global_feat=True)

    def _init_pose_predictor(self):
        dim = self.pc_feat_dim * 2
        if self.semantic:
            dim += self.max_num_part
        if self.use_part_label:
            dim += self.cfg.data.num_part_category
        return StocasticPoseRegressor(feat_dim=dim, noise_dim=self.cfg.loss.noise_dim,
                                      rot_type=self.rot_type)

    def _extract_part_feats(self, part_pcs, part_valids):
        return kernels.encode_parts(self.encoder, part_pcs, part_valids, self.pc_feat_dim)

    def _extract_global_feats(self, part_pcs):
        return self.global_encoder(part_pcs.flatten(1, 2))  # [B, C]

    def forward(self, data_dict):
        feats = data_dict.get('pre_pose_feats', None)
        if feats is None:
            part_pcs = data_dict['part_pcs']
            pc_feats = self._extract_part_feats(part_pcs, data_dict['part_valids'])
            global_feats = self._extract_global_feats(part_pcs)
            global_feats = global_feats.unsqueeze(1).expand(-1, self.max_num_part, -1)
            feats = torch.cat([global_feats, pc_feats, data_dict['part_label'].type_as(pc_feats),
                               data_dict['instance_label'].type_as(pc_feats)], dim=-1)
        rot, trans = self.pose_predictor(feats)
        return {'rot': self._wrap_rotation(rot), 'trans': trans, 'pre_pose_feats': feats}

    def _loss_function(self, data_dict, out_dict={}, optimizer_idx=-1):
        forward_dict = {k: data_dict[k] for k in
                        ('part_pcs', 'part_valids', 'part_label', 'instance_label')}
        forward_dict['pre_pose_feats'] = out_dict.get('pre_pose_feats', None)
        pred = self.forward(forward_dict)
        loss_dict, out_dict = self._calc_loss(pred, data_dict)
        out_dict['pre_pose_feats'] = pred['pre_pose_feats']
        return loss_dict, out_dict
