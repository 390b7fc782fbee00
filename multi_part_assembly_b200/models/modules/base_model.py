"""`BaseModel`: the LightningModule-style base class of every assembly model,
same hook names and semantics as the reference
(models/modules/base_model.py:17-464): `forward`, `training_step`,
`validation_step`/`test_step` + `*_epoch_end`, `forward_pass`, `_calc_loss`,
`loss_function` (Min-of-N), `_match_parts` (Hungarian matching of
geometrically equivalent parts), `configure_optimizers`.

Differences that do not change results: the loss terms run on the fused CUDA
ops of utils/loss.py, per-step logging keeps tensors instead of calling
`.item()` on every term (a host sync per term in the reference, :138), and the
matching cost matrices of all groups of a batch are solved by one batched
assignment kernel on the device (no device->host copy of the costs).
"""
import numpy as np
import torch
import torch.optim as optim

from ...compat.lightning import LightningModule
from ...utils import transform_pc, Rotation3D, filter_wd_parameters
from ...utils import trans_l2_loss, rot_points_cd_loss, shape_cd_loss, \
    rot_cosine_loss, rot_points_l2_loss, chamfer_distance
from ...utils import calc_part_acc, calc_connectivity_acc, trans_metrics, \
    rot_metrics
from ...utils.loss import fused_geometric_losses
from ...utils.lr import CosineAnnealingWarmupRestarts


class BaseModel(LightningModule):
    """Base class for multi-part assembly models."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self._setup()

    def _setup(self):
        self.rot_type = self.cfg.model.rot_type
        if self.rot_type == 'quat':
            self.pose_dim = 3 + 4
            ones = (0, )
        elif self.rot_type == 'rmat':
            self.pose_dim = 3 + 6
            ones = (0, 4)
        else:
            raise NotImplementedError(f'rotation {self.rot_type} is not supported')
        zero_pose = torch.zeros(1, 1, self.pose_dim)
        for i in ones:
            zero_pose[..., i] = 1.
        self.zero_pose = zero_pose  # plain attribute like the reference (not in the state dict)
        self._zero_pose_cache = {}

        self.semantic = (self.cfg.data.dataset != 'geometry')
        self.max_num_part = self.cfg.data.max_num_part
        self.pc_feat_dim = self.cfg.model.pc_feat_dim
        self.use_part_label = 'part_label' in self.cfg.data.data_keys
        self.sample_iter = self.cfg.loss.get('sample_iter', 1)

    # ------------------------------------------------------------------ hooks
    def forward(self, data_dict):
        """Predict poses for each part (implemented by the subclasses)."""
        raise NotImplementedError

    def _zero_pose_like(self, ref, B, P):
        """[B, P, pose_dim] identity poses on ref's device/dtype.  The reference moves the
        host constant with `.type_as` on every forward (a pageable host-to-device copy, which
        also cannot be captured into a CUDA graph); here the device copy is made once."""
        key = (ref.device, ref.dtype)
        z = self._zero_pose_cache.get(key)
        if z is None:
            z = self._zero_pose_cache[key] = self.zero_pose.to(device=ref.device, dtype=ref.dtype)
        return z.repeat(B, P, 1).detach()

    def training_step(self, data_dict, batch_idx, optimizer_idx=-1):
        return self.forward_pass(data_dict, mode='train', optimizer_idx=optimizer_idx)['loss']

    def validation_step(self, data_dict, batch_idx):
        return self.forward_pass(data_dict, mode='val', optimizer_idx=-1)

    def test_step(self, data_dict, batch_idx):
        return self.forward_pass(data_dict, mode='test', optimizer_idx=-1)

    @staticmethod
    def _weighted_epoch_mean(outputs, prefix, stack_loss=True):
        int_bs = isinstance(outputs[0]['batch_size'], int)
        bs_fn = torch.tensor if int_bs else (torch.stack if stack_loss else torch.cat)
        loss_fn = torch.stack if (int_bs or stack_loss) else torch.cat
        batch_sizes = bs_fn([o.pop('batch_size') for o in outputs]).type_as(outputs[0]['loss'])
        return {
            f'{prefix}/{k}': (loss_fn([o[k] for o in outputs]) * batch_sizes).sum() / batch_sizes.sum()
            for k in outputs[0].keys()
        }

    def validation_epoch_end(self, outputs):
        self.log_dict(self._weighted_epoch_mean(outputs, 'val'), sync_dist=True)

    def test_epoch_end(self, outputs):
        avg_loss = self._weighted_epoch_mean(outputs, 'test', stack_loss=False)
        print('; '.join([f'{k}: {v.item():.6f}' for k, v in avg_loss.items()]))
        self.test_results = avg_loss

    def forward_pass(self, data_dict, mode, optimizer_idx):
        """Loss computation and logging for one batch.  `data_dict` follows the
        dataset schema; like the reference (:130-132)
        the call replaces `part_quat` by the Rotation3D `part_rot` in place."""
        part_quat = data_dict.pop('part_quat')
        part_rot = Rotation3D(part_quat, rot_type='quat')  # a fresh tensor: no defensive clone
        data_dict['part_rot'] = part_rot if self.rot_type == 'quat' else part_rot.convert(self.rot_type)
        loss_dict = self.loss_function(data_dict, optimizer_idx=optimizer_idx)

        if mode == 'train' and self.local_rank == 0:
            log_dict = {f'{mode}/{k}': v.detach() for k, v in loss_dict.items()}
            profiler = getattr(getattr(self, 'trainer', None), 'profiler', None)
            if profiler is not None:
                names = [k for k in profiler.recorded_durations if 'prepare_data' in k]
                if names:
                    log_dict[f'{mode}/data_time'] = profiler.recorded_durations[names[0]][-1]
            self.log_dict(log_dict, logger=True, sync_dist=False, rank_zero_only=True)
        return loss_dict

    # --------------------------------------------------------------- matching
    @torch.no_grad()
    def _match_cost(self, pts, trans1, rot1, trans2, rot2):
        """p x p matching cost of one group of equivalent parts: Chamfer
        distance between part i under pose-1 i and part j ... (reference
        :162-174).  Returns the [p, p] device tensor."""
        p, N, _ = pts.shape
        n = 100
        sample_idx = torch.randperm(N)[:n].to(pts.device).long()  # CPU RNG, as the reference
        pts = pts[:, sample_idx]
        pts1 = transform_pc(trans1, rot1, pts, self.rot_type)
        pts2 = transform_pc(trans2, rot2, pts, self.rot_type)
        pts1 = pts1.unsqueeze(1).expand(p, p, n, 3).reshape(-1, n, 3)
        pts2 = pts2.unsqueeze(0).expand(p, p, n, 3).reshape(-1, n, 3)
        dist1, dist2 = chamfer_distance(pts1, pts2)
        return (dist1.mean(1) + dist2.mean(1)).view(p, p)

    @torch.no_grad()
    def _linear_sum_assignment(self, pts, trans1, rot1, trans2, rot2):
        """Min-cost matching between two groups of poses (reference :150-179)."""
        from ... import kernels
        dist_mat = self._match_cost(pts, trans1, rot1, trans2, rot2)
        cind = kernels.lsap_batched([dist_mat])[0]
        return torch.arange(cind.shape[0], device=cind.device), cind

    def _match_groups(self, match_ids):
        """Groups of equivalent parts of a batch, in the reference's visiting order (shape by
        shape, group id ascending, :207-218): list of (shape index, member part indices).
        The group structure comes from the data loader; it is read on the host once per
        batch (the Min-of-N samples of a step reuse it) because the reference draws one CPU
        `torch.randperm` per group (:165) -- the number of draws is data dependent."""
        key = (match_ids.data_ptr(), match_ids._version, tuple(match_ids.shape))
        cache = getattr(self, '_match_cache', None)
        if cache is not None and cache[0] == key:
            return cache[1]
        ids_host = match_ids.long().cpu().numpy()
        groups = []
        for ind in range(ids_host.shape[0]):
            for g in range(1, int(ids_host[ind].max()) + 1):
                members = np.nonzero(ids_host[ind] == g)[0]
                if len(members):
                    groups.append((ind, members.astype(np.int32)))
                else:  # an empty group still costs the reference one randperm draw
                    groups.append((ind, members.astype(np.int32)))
        self._match_cache = (key, groups)
        return groups

    @torch.no_grad()
    def _match_parts(self, part_pcs, pred_trans, pred_rot, gt_trans, gt_rot, match_ids):
        """Semantic assembly: permute the GT poses inside every group of
        geometrically equivalent parts to the min-cost assignment against the
        predictions (reference :181-238).  `match_ids` [B, P]: 0 = unique or
        padded, g > 0 = group id.

        All groups of the batch go through ONE fused call (csrc/loss.cu `mpa_match_parts`:
        cost matrices, assignments, permutation -- three launches, nothing returns to the
        host); the host only builds the group table and draws the subsamples with the
        reference's RNG calls."""
        from ... import _lib
        groups = self._match_groups(match_ids)
        new_gt_trans = gt_trans.detach().clone().float().contiguous()
        new_gt_quat = gt_rot.rot.detach().clone().float().contiguous()
        B, P, N, _ = part_pcs.shape
        n = min(100, N)
        if self.rot_type != 'quat' or not part_pcs.is_cuda:
            return self._match_parts_loop(part_pcs, pred_trans, pred_rot, gt_trans, gt_rot, groups)
        # the reference's RNG consumption: one randperm(N) per group, in visiting order
        samples = [torch.randperm(N)[:n] for _ in groups]
        live = [(k, ind, m) for k, (ind, m) in enumerate(groups) if len(m) > 0]
        if not live:
            return new_gt_trans.type_as(gt_trans), self._wrap_rotation(new_gt_quat.type_as(gt_rot.rot))
        G = len(live)
        MAXP = 32
        sizes = np.array([len(m) for _, _, m in live], np.int32)
        assert sizes.max() <= MAXP, 'groups of at most 32 equivalent parts'
        cost_off = np.concatenate(([0], np.cumsum(sizes.astype(np.int64)**2)[:-1])).astype(np.int32)
        out_off = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)
        members = np.zeros((G, MAXP), np.int32)
        for r, (_, _, m) in enumerate(live):
            members[r, :len(m)] = m
        sample = torch.stack([samples[k] for k, _, _ in live]).to(torch.int32).numpy()
        pairs = np.concatenate([(r << 16) | (np.arange(p)[:, None] << 8) | np.arange(p)[None, :]
                                for r, p in enumerate(sizes)], axis=None).astype(np.int32)
        table = np.concatenate([np.array([ind for _, ind, _ in live], np.int32), sizes, cost_off, out_off,
                                members.ravel(), sample.ravel(), pairs])
        dev = part_pcs.device
        table_d = torch.from_numpy(table).to(dev, non_blocking=True)
        n_pairs, total_rows = int(pairs.size), int(sizes.sum())
        L = _lib.lib()
        pts = part_pcs.detach().float().contiguous()
        q1 = pred_rot.rot.detach().float().contiguous()
        t1 = pred_trans.detach().float().contiguous()
        q2 = gt_rot.rot.detach().float().contiguous()
        t2 = gt_trans.detach().float().contiguous()
        ws_bytes = L.mpa_match_parts_workspace_bytes(n_pairs, total_rows)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = L.mpa_match_parts(_lib.ptr(pts), _lib.ptr(q1), _lib.ptr(t1), _lib.ptr(q2), _lib.ptr(t2),
                                   B, P, N, n, _lib.ptr(table_d), G, n_pairs, total_rows, int(sizes.max()),
                                   _lib.ptr(new_gt_trans), _lib.ptr(new_gt_quat), None, None,
                                   _lib.ptr(ws), ws_bytes, _lib.cuda_stream(dev))
        _lib.check(rc, 'mpa_match_parts')
        return new_gt_trans.type_as(gt_trans), self._wrap_rotation(new_gt_quat.type_as(gt_rot.rot))

    @torch.no_grad()
    def _match_parts_loop(self, part_pcs, pred_trans, pred_rot, gt_trans, gt_rot, groups):
        """Per-group formulation (rotation-matrix poses): cost matrices with the generic ops,
        all assignments in one launch."""
        new_gt_trans = gt_trans.detach().clone()
        new_gt_rot_tensor = gt_rot.detach().clone().rot
        gt_rot_tensor = gt_rot.rot
        pred_rot_tensor = pred_rot.rot
        dev = part_pcs.device
        groups = [(ind, torch.from_numpy(m.astype(np.int64)).to(dev)) for ind, m in groups]
        costs = [self._match_cost(part_pcs[ind, m], pred_trans[ind, m], pred_rot_tensor[ind, m],
                                  gt_trans[ind, m], new_gt_rot_tensor[ind, m])
                 for ind, m in groups]
        if costs:
            from ... import kernels
            for (ind, m), cind in zip(groups, kernels.lsap_batched(costs)):
                new_gt_trans[ind, m] = gt_trans[ind, m][cind]
                new_gt_rot_tensor[ind, m] = gt_rot_tensor[ind, m][cind]
        return new_gt_trans, self._wrap_rotation(new_gt_rot_tensor)

    # ------------------------------------------------------------------- loss
    def _calc_loss(self, out_dict, data_dict):
        """All loss terms ([B] each) for one prediction; evaluation metrics too
        when not training (reference :240-314)."""
        # the single-kernel reduction of loss_function is only valid when ONE prediction
        # was scored (DGL / refine models score every GNN / refine iteration)
        self._calc_loss_calls = getattr(self, '_calc_loss_calls', 0) + 1
        pred_trans, pred_rot = out_dict['trans'], out_dict['rot']
        # last scored prediction (references, no copies): lets a CUDA-graph replay expose
        # the poses next to the losses (runtime.GraphedStep.static_pred)
        self._last_pred = (pred_trans, pred_rot.rot)
        part_pcs, valids = data_dict['part_pcs'], data_dict['part_valids']
        gt_trans, gt_rot = data_dict['part_trans'], data_dict['part_rot']
        if self.semantic:
            new_trans, new_rot = self._match_parts(part_pcs, pred_trans, pred_rot, gt_trans,
                                                   gt_rot, data_dict['match_ids'])
        elif self._can_fuse_losses(pred_rot, gt_rot):
            # the fused kernels only read the ground-truth poses: no copies
            return self._calc_loss_fused(out_dict, data_dict, gt_trans, gt_rot)
        else:
            new_trans, new_rot = gt_trans.detach().clone(), gt_rot.detach().clone()

        if self._can_fuse_losses(pred_rot, new_rot):
            return self._calc_loss_fused(out_dict, data_dict, new_trans, new_rot)

        trans_loss = trans_l2_loss(pred_trans, new_trans, valids)
        rot_pt_cd_loss = rot_points_cd_loss(part_pcs, pred_rot, new_rot, valids)
        transform_pt_cd_loss, pred_trans_pts, gt_trans_pts = shape_cd_loss(
            part_pcs, pred_trans, new_trans, pred_rot, new_rot, valids, ret_pts=True,
            # semantic: always divide by the padded part count; geometric: only
            # while training (reference :264-280)
            training=self.semantic or self.training)
        loss_dict = {
            'trans_loss': trans_loss,
            'rot_pt_cd_loss': rot_pt_cd_loss,
            'transform_pt_cd_loss': transform_pt_cd_loss,
        }
        if self.cfg.loss.use_rot_loss:
            loss_dict['rot_loss'] = rot_cosine_loss(pred_rot, new_rot, valids)
        if self.cfg.loss.use_rot_pt_l2_loss:
            loss_dict['rot_pt_l2_loss'] = rot_points_l2_loss(part_pcs, pred_rot, new_rot, valids)
        if not self.training:
            loss_dict.update(self._calc_metrics(data_dict, out_dict, new_trans, new_rot))
        out_dict = {
            'pred_trans': pred_trans,
            'pred_rot': pred_rot,
            'gt_trans_pts': gt_trans_pts,
            'pred_trans_pts': pred_trans_pts,
        }
        return loss_dict, out_dict

    def _can_fuse_losses(self, pred_rot, gt_rot):
        """The fused forward-only loss kernels apply when nothing needs a
        gradient (evaluation, benchmarking, CUDA-graph replay) and both
        rotations are quaternions."""
        return (not torch.is_grad_enabled()) and pred_rot.rot_type == 'quat' and \
            gt_rot.rot_type == 'quat' and pred_rot.rot.is_cuda

    def _loss_weight_tensor(self, device):
        cache = getattr(self, '_loss_w_cache', None)
        if cache is None or cache.device != device:
            c = self.cfg.loss
            w = [c.trans_loss_w, c.rot_pt_cd_loss_w, c.transform_pt_cd_loss_w,
                 c.rot_loss_w if c.use_rot_loss else 0.,
                 c.rot_pt_l2_loss_w if c.use_rot_pt_l2_loss else 0.]
            cache = torch.tensor(w, dtype=torch.float32, device=device)
            self._loss_w_cache = cache
        return cache

    def _calc_loss_fused(self, out_dict, data_dict, new_trans, new_rot):
        pred_trans, pred_rot = out_dict['trans'], out_dict['rot']
        part_pcs, valids = data_dict['part_pcs'], data_dict['part_valids']
        terms, pred_trans_pts, gt_trans_pts = fused_geometric_losses(
            part_pcs, pred_trans, new_trans, pred_rot, new_rot, valids,
            self._loss_weight_tensor(part_pcs.device), training=self.semantic or self.training,
            want_rot_l2=bool(self.cfg.loss.use_rot_pt_l2_loss), ret_pts=True)
        loss_dict = {k: terms[k] for k in ('trans_loss', 'rot_pt_cd_loss', 'transform_pt_cd_loss')}
        if self.cfg.loss.use_rot_loss:
            loss_dict['rot_loss'] = terms['rot_loss']
        if self.cfg.loss.use_rot_pt_l2_loss:
            loss_dict['rot_pt_l2_loss'] = terms['rot_pt_l2_loss']
        # the packed [6, B] tensor (terms + weighted total) lets loss_function reduce
        # everything with one kernel when there is a single sample
        self._fused_packed = (terms.packed, tuple(loss_dict.keys()))
        if not self.training:
            loss_dict.update(self._calc_metrics(data_dict, out_dict, new_trans, new_rot))
        return loss_dict, {'pred_trans': pred_trans, 'pred_rot': pred_rot,
                           'gt_trans_pts': gt_trans_pts, 'pred_trans_pts': pred_trans_pts}

    @torch.no_grad()
    def _calc_metrics(self, data_dict, out_dict, gt_trans, gt_rot):
        """Evaluation metrics (reference :316-339)."""
        metric_dict = {}
        part_pcs, valids = data_dict['part_pcs'], data_dict['part_valids']
        pred_trans, pred_rot = out_dict['trans'], out_dict['rot']
        metric_dict['part_acc'] = calc_part_acc(part_pcs, pred_trans, gt_trans, pred_rot,
                                                gt_rot, valids)
        if self.semantic and 'contact_points' in data_dict.keys():
            metric_dict['connectivity_acc'] = calc_connectivity_acc(
                pred_trans, pred_rot, data_dict['contact_points'])
        if not self.semantic:
            for metric in ['mse', 'rmse', 'mae']:
                metric_dict[f'trans_{metric}'] = trans_metrics(pred_trans, gt_trans, valids,
                                                               metric=metric)
                metric_dict[f'rot_{metric}'] = rot_metrics(pred_rot, gt_rot, valids, metric=metric)
        return metric_dict

    def _loss_function(self, data_dict, out_dict={}, optimizer_idx=-1):
        raise NotImplementedError

    def loss_function(self, data_dict, optimizer_idx):
        """Min-of-N loss: sample `sample_iter` predictions, keep per shape the
        one with the lowest weighted total (reference :348-387)."""
        samples = None
        out_dict = {}
        self._fused_packed = None
        self._calc_loss_calls = 0
        for _ in range(self.sample_iter):
            sample_loss, out_dict = self._loss_function(data_dict, out_dict,
                                                        optimizer_idx=optimizer_idx)
            if samples is None:
                samples = {k: [] for k in sample_loss.keys()}
            for k, v in sample_loss.items():
                samples[k].append(v)
        packed = self._fused_packed
        if self.sample_iter == 1 and packed is not None and self._calc_loss_calls == 1 and \
                tuple(k for k in samples.keys() if k.endswith('_loss')) == packed[1]:
            # single sample straight from the fused loss kernels: one mean over [6, B]
            terms, keys = packed
            means = terms.mean(dim=1)
            order = ('trans_loss', 'rot_pt_cd_loss', 'transform_pt_cd_loss', 'rot_loss',
                     'rot_pt_l2_loss')
            loss_dict = {k: means[order.index(k)] for k in keys}
            for k, v in samples.items():  # evaluation metrics computed next to the loss terms
                if k not in loss_dict:
                    loss_dict[k] = v[0].mean()
            loss_dict['loss'] = means[5]
            if not self.training:
                loss_dict['batch_size'] = terms.shape[1]
            return loss_dict
        loss_dict = {k: torch.stack(v, dim=0) for k, v in samples.items()}
        total_loss = 0.
        for k, v in loss_dict.items():
            if k.endswith('_loss'):  # metrics logged in eval are not part of the loss
                total_loss = total_loss + v * self.cfg.loss[f'{k}_w']
        loss_dict['loss'] = total_loss
        if self.sample_iter == 1:
            loss_dict = {k: v[0].mean() for k, v in loss_dict.items()}
        else:
            min_idx = total_loss.argmin(0)  # [B]
            batch_idx = torch.arange(min_idx.shape[0]).type_as(min_idx)
            loss_dict = {k: v[min_idx, batch_idx].mean() for k, v in loss_dict.items()}
        if not self.training:
            loss_dict['batch_size'] = total_loss.shape[1]
        return loss_dict

    # -------------------------------------------------------------- optimiser
    def configure_optimizers(self):
        """Adam(W) + optional cosine schedule with warm-up (reference :389-425)."""
        lr = self.cfg.optimizer.lr
        wd = self.cfg.optimizer.weight_decay
        # same update rule as the reference; on CUDA parameters the multi-tensor ("fused")
        # implementation applies it in one launch instead of a few hundred
        extra = {'fused': True} if all(p.is_cuda for p in self.parameters()) else {}
        if wd > 0.:
            groups = filter_wd_parameters(self)
            optimizer = optim.AdamW([
                {'params': groups['no_decay'], 'weight_decay': 0.},
                {'params': groups['decay'], 'weight_decay': wd},
            ], lr=lr, **extra)
        else:
            optimizer = optim.Adam(self.parameters(), lr=lr, weight_decay=0., **extra)
        if self.cfg.optimizer.lr_scheduler:
            assert self.cfg.optimizer.lr_scheduler in ['cosine']
            total_epochs = self.cfg.exp.num_epochs
            warmup_epochs = int(total_epochs * self.cfg.optimizer.warmup_ratio)
            scheduler = CosineAnnealingWarmupRestarts(
                optimizer, total_epochs, max_lr=lr,
                min_lr=lr / self.cfg.optimizer.lr_decay_factor, warmup_steps=warmup_epochs)
            return [optimizer], [{'scheduler': scheduler, 'interval': 'epoch'}]
        return optimizer

    @torch.no_grad()
    def sample_assembly(self, data_dict):
        """Predicted and GT assembled clouds per shape (reference :427-460),
        as lists of [p*N, 3] arrays (colouring is left to the caller)."""
        if 'part_rot' not in data_dict:
            part_quat = data_dict.pop('part_quat')
            data_dict['part_rot'] = Rotation3D(part_quat, rot_type='quat').convert(self.rot_type)
        part_pcs, valids = data_dict['part_pcs'], data_dict['part_valids']
        gt_pcs = transform_pc(data_dict['part_trans'], data_dict['part_rot'], part_pcs)
        B = part_pcs.shape[0]
        pred_pcs_lst, gt_pcs_lst = [[] for _ in range(B)], []
        for i in range(self.sample_iter):
            out = self.forward(data_dict)
            pred_pcs = transform_pc(out['trans'], out['rot'], part_pcs)
            for j in range(B):
                valid = valids[j].bool()
                pred_pcs_lst[j].append(pred_pcs[j][valid].cpu().numpy().reshape(-1, 3))
                if i == 0:
                    gt_pcs_lst.append(gt_pcs[j][valid].cpu().numpy().reshape(-1, 3))
        return gt_pcs_lst, pred_pcs_lst

    def _wrap_rotation(self, rot_tensor):
        return Rotation3D(rot_tensor, rot_type=self.rot_type)
