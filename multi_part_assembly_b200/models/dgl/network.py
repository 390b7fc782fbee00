"""Dynamic graph learning assembly model, `DGL` (reference
models/dgl/network.py:14-297): part encoder, `gnn_iter` rounds of dense P x P
edge-MLP message passing with learned relation weights, one pose head per
round.  The GNN MLPs stay stock torch modules (the encoder
and the per-iteration Chamfer losses are the hot part)."""
import numpy as np
import torch

from ..modules.base_model import BaseModel
from ..modules.encoder import build_encoder
from ..modules.regressor import StocasticPoseRegressor
from ...utils import _get_clones
from ... import kernels
from .modules import MLP3, MLP4, RelationNet, PoseEncoder


class DGLModel(BaseModel):

    def __init__(self, cfg):
        super().__init__(cfg)
        self.iter = self.cfg.model.gnn_iter
        self.merge_node = self.cfg.model.merge_node
        self.encoder = build_encoder(self.cfg.model.encoder, feat_dim=self.pc_feat_dim,
                                     global_feat=True)
        self.edge_mlps = _get_clones(MLP3(self.pc_feat_dim), self.iter)
        self.node_mlps = _get_clones(MLP4(self.pc_feat_dim), self.iter)
        self.pose_predictors = self._init_pose_predictor()
        self.relation_predictor_dense = RelationNet()
        if self.merge_node:
            self.relation_predictor = RelationNet()
        self.pose_extractor = PoseEncoder(self.pose_dim)

    def _init_pose_predictor(self):
        dim = self.pc_feat_dim + self.pose_dim
        if self.semantic:
            dim += self.max_num_part
        if self.use_part_label:
            dim += self.cfg.data.num_part_category
        head = StocasticPoseRegressor(feat_dim=dim, noise_dim=self.cfg.loss.noise_dim,
                                      rot_type=self.rot_type)
        return _get_clones(head, self.iter)

    def _gather_same_class(self, data_dict):
        """Per shape, the index lists of valid parts sharing a part id."""
        class_list = data_dict.get('class_list', None)
        if self.merge_node and self.semantic and class_list is None:
            valids = data_dict['part_valids'].cpu().numpy()
            ids = data_dict['part_ids'].cpu().numpy()
            class_list = []
            for i in range(valids.shape[0]):
                class_ids = ids[i][valids[i] == 1]
                class_list.append([np.where(class_ids == lbl)[0] for lbl in np.unique(class_ids)])
        return class_list

    def _extract_part_feats(self, part_pcs, part_valids):
        return kernels.encode_parts(self.encoder, part_pcs, part_valids, self.pc_feat_dim)

    def _merge_nodes(self, part_feats, pose_feats, class_list):
        """Max-pool features over each class of equivalent parts (:101-119)."""
        pose_out, part_out = pose_feats.clone(), part_feats.clone()
        for i in range(part_feats.shape[0]):
            for cls_lst in class_list[i]:
                if len(cls_lst) <= 1:
                    continue
                pose_out[i, cls_lst] = pose_feats[i, cls_lst].max(dim=-2, keepdim=True)[0]
                part_out[i, cls_lst] = part_feats[i, cls_lst].max(dim=-2, keepdim=True)[0]
        return part_out, pose_out

    def _update_relation(self, pose_feats, iter_ind):
        """[B, P, F] pose features -> [B, P, P] relation weights (:121-133)."""
        B, P, _ = pose_feats.shape
        pair = torch.cat([pose_feats.unsqueeze(1).expand(B, P, P, -1),
                          pose_feats.unsqueeze(2).expand(B, P, P, -1)], dim=-1)
        net = self.relation_predictor if (self.merge_node and iter_ind % 2 == 1) \
            else self.relation_predictor_dense
        return net(pair.reshape(B, P * P, -1)).view(B, P, P)

    def _message_passing(self, part_feats, relation_matrix, iter_ind):
        """Relation-weighted mean of the edge-MLP outputs (:135-152)."""
        B, P, _ = part_feats.shape
        pair = torch.cat([part_feats.unsqueeze(2).expand(B, P, P, -1),
                          part_feats.unsqueeze(1).expand(B, P, P, -1)], dim=-1)
        edge = self.edge_mlps[iter_ind](pair.reshape(B * P, P, -1)).view(B, P, P, -1)
        message = (edge * relation_matrix.unsqueeze(-1)).sum(dim=2)
        return message / (relation_matrix.sum(dim=-1, keepdim=True) + 1e-6)

    def forward(self, data_dict):
        part_feats = data_dict.get('part_feats', None)
        if part_feats is None:
            part_feats = self._extract_part_feats(data_dict['part_pcs'], data_dict['part_valids'])
        local_feats = part_feats
        valid_matrix = data_dict['valid_matrix']
        part_label = data_dict['part_label'].type_as(part_feats)
        instance_label = data_dict['instance_label'].type_as(part_feats)
        B, P = instance_label.shape[:2]
        pred_pose = self._zero_pose_like(part_feats, B, P)
        class_list = self._gather_same_class(data_dict)

        all_rot, all_trans = [], []
        for it in range(self.iter):
            if it >= 1:
                pose_feats = self.pose_extractor(pred_pose)
                if self.merge_node and self.semantic and it % 2 == 1:
                    feats_in, pose_in = self._merge_nodes(part_feats, pose_feats, class_list)
                else:
                    feats_in, pose_in = part_feats, pose_feats
                relation = self._update_relation(pose_in, it) * valid_matrix
            else:
                feats_in, relation = part_feats, valid_matrix
            messages = self._message_passing(feats_in, relation, it)
            node_feats = torch.cat([messages.type_as(part_feats), part_feats], dim=-1)
            part_feats = self.node_mlps[it](node_feats)
            pose_feats = torch.cat([part_feats, part_label, instance_label, pred_pose], dim=-1)
            rot, trans = self.pose_predictors[it](pose_feats)
            pred_pose = torch.cat([rot, trans], dim=-1)
            all_rot.append(rot)
            all_trans.append(trans)

        if self.training:
            rot = self._wrap_rotation(torch.stack(all_rot, dim=0))
            trans = torch.stack(all_trans, dim=0)
        else:
            rot = self._wrap_rotation(all_rot[-1])
            trans = all_trans[-1]
        return {'rot': rot, 'trans': trans, 'part_feats': local_feats, 'class_list': class_list}

    def _loss_function(self, data_dict, out_dict={}, optimizer_idx=-1):
        """Training: the loss of every GNN iteration is summed (:284-297)."""
        forward_dict = {k: data_dict[k] for k in
                        ('part_pcs', 'part_valids', 'part_label', 'instance_label', 'part_ids',
                         'valid_matrix')}
        forward_dict['part_feats'] = out_dict.get('part_feats', None)
        forward_dict['class_list'] = out_dict.get('class_list', None)
        pred = self.forward(forward_dict)
        part_feats, class_list = pred['part_feats'], pred['class_list']
        if not self.training:
            loss_dict, out_dict = self._calc_loss(pred, data_dict)
        else:
            loss_dict = None
            for i in range(self.iter):
                step_loss, out_dict = self._calc_loss(
                    {'rot': pred['rot'][i], 'trans': pred['trans'][i]}, data_dict)
                if loss_dict is None:
                    loss_dict = {k: 0. for k in step_loss.keys()}
                for k, v in step_loss.items():
                    loss_dict[k] = loss_dict[k] + v
                    loss_dict[f'{k}_{i}'] = v
        out_dict['part_feats'] = part_feats
        out_dict['class_list'] = class_list
        return loss_dict, out_dict
