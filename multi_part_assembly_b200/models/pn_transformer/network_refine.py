"""PNTransformer with iterative pose refinement (reference
models/pn_transformer/network_refine.py:11-175)."""
import torch
import torch.nn as nn

from ..modules.regressor import StocasticPoseRegressor
from ...utils import _get_clones
from .network import PNTransformer
from .transformer import TransformerEncoder


class PosEncoder(nn.Module):
    """MLP positional encoding of the current pose."""

    def __init__(self, dims):
        super().__init__()
        layers = []
        for i in range(len(dims) - 2):
            layers += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU()]
        layers.append(nn.Linear(dims[-2], dims[-1]))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)


class PNTransformerRefine(PNTransformer):

    def __init__(self, cfg):
        self.refine_steps = cfg.model.refine_steps
        self.pose_pc_feat = cfg.model.pose_pc_feat
        super().__init__(cfg)
        self.corr_pos_enc = PosEncoder([self.pose_dim] + list(self.cfg.model.transformer_pos_enc))

    def _init_corr_module(self):
        m = self.cfg.model
        corr = TransformerEncoder(
            d_model=self.pc_feat_dim, num_heads=m.transformer_heads,
            ffn_dim=m.transformer_feat_dim, num_layers=m.transformer_layers,
            norm_first=m.transformer_pre_ln, out_dim=self.pc_feat_dim)
        return _get_clones(corr, self.refine_steps)

    def _init_pose_predictor(self):
        dim = self._pose_in_dim() + self.pose_dim
        if self.pose_pc_feat:
            dim += self.pc_feat_dim
        head = StocasticPoseRegressor(feat_dim=dim, noise_dim=self.cfg.loss.noise_dim,
                                      rot_type=self.rot_type)
        return _get_clones(head, self.refine_steps)

    def forward(self, data_dict):
        pc_feats = data_dict.get('pc_feats', None)
        part_pcs, part_valids = data_dict['part_pcs'], data_dict['part_valids']
        if pc_feats is None:
            pc_feats = self._extract_part_feats(part_pcs, part_valids)
        part_feats = pc_feats
        part_label = data_dict['part_label'].type_as(pc_feats)
        inst_label = data_dict['instance_label'].type_as(pc_feats)
        B, P, _ = inst_label.shape
        pose = self._zero_pose_like(part_feats, B, P)
        valid_mask = part_valids == 1
        pred_rot, pred_trans = [], []
        for i in range(self.refine_steps):
            in_feats = part_feats + self.corr_pos_enc(pose)
            corr_feats = self.corr_module[i](in_feats, valid_mask)
            feats = torch.cat([corr_feats, part_label, inst_label, pose], dim=-1)
            if self.pose_pc_feat:
                feats = torch.cat([pc_feats, feats], dim=-1)
            rot, trans = self.pose_predictor[i](feats)
            pred_rot.append(rot)
            pred_trans.append(trans)
            part_feats = corr_feats
            pose = torch.cat([rot, trans], dim=-1)
        if self.training:
            rot = self._wrap_rotation(torch.stack(pred_rot, dim=0))
            trans = torch.stack(pred_trans, dim=0)
        else:
            rot = self._wrap_rotation(pred_rot[-1])
            trans = pred_trans[-1]
        return {'rot': rot, 'trans': trans, 'pc_feats': pc_feats}

    def _loss_function(self, data_dict, out_dict={}, optimizer_idx=-1):
        forward_dict = {k: data_dict[k] for k in
                        ('part_pcs', 'part_valids', 'part_label', 'instance_label')}
        forward_dict['pc_feats'] = out_dict.get('pc_feats', None)
        pred = self.forward(forward_dict)
        pc_feats = pred['pc_feats']
        if not self.training:
            loss_dict, out_dict = self._calc_loss(pred, data_dict)
            out_dict['pc_feats'] = pc_feats
            return loss_dict, out_dict
        all_loss = None
        for i in range(self.refine_steps):
            loss_dict, out_dict = self._calc_loss(
                {'rot': pred['rot'][i], 'trans': pred['trans'][i]}, data_dict)
            if all_loss is None:
                all_loss = {k: 0. for k in loss_dict.keys()}
            for k, v in loss_dict.items():
                all_loss[k] = all_loss[k] + v
                all_loss[f'{k}_{i}'] = v
        out_dict['pc_feats'] = pc_feats
        return all_loss, out_dict
