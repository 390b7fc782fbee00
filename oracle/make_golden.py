"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python
(from /root/reference, under oracle/ref_shims.py) on seeded inputs.

    python -m oracle.make_golden        # build container only

Weights are not stored: both sides fill them with oracle/params.fill_params_.
Dropout is set to p=0 on the reference side (SURVEY.md 8c parity hazard 1).
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tests', 'golden')

from oracle import ref_shims  # noqa: E402
from oracle.params import fill_params_  # noqa: E402
from multi_part_assembly_b200.datasets.synthetic import make_batch  # noqa: E402
from multi_part_assembly_b200.compat import lightning  # noqa: E402


def npz(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    clean = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        clean[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **clean)
    print(f'{name}: ' + ', '.join(f'{k}{list(v.shape)}' for k, v in clean.items()))


def zero_dropout(module):
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    return module


def load_ref_cfg(rel):
    import importlib.util
    path = os.path.join(ref_shims.REFERENCE_ROOT, 'configs', rel)
    spec = importlib.util.spec_from_file_location('golden_cfg_' + str(abs(hash(rel))), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_cfg_defaults()


def main():
    ref = ref_shims.install()
    torch.set_num_threads(8)
    from multi_part_assembly.utils import (transform_pc, rot_pc, Rotation3D, trans_l2_loss,
                                           rot_cosine_loss, rot_points_l2_loss, rot_points_cd_loss,
                                           shape_cd_loss, calc_part_acc, trans_metrics, rot_metrics,
                                           rot_l2_loss, chamfer_distance, calc_connectivity_acc)
    from multi_part_assembly.models import build_model, build_encoder, StocasticPoseRegressor
    from multi_part_assembly.models.modules.encoder.dgcnn import knn
    from multi_part_assembly.models.pn_transformer.transformer import TransformerEncoder

    # 1. Chamfer, BASELINE config A, against the reference test's own definition
    _, nn_distance_torch = ref_shims.load_reference_test_functions()
    torch.manual_seed(0)
    x1, x2 = torch.rand(2, 500, 3), torch.rand(2, 500, 3)
    d1, i1, d2, i2 = nn_distance_torch(x1, x2, 'NWC')
    npz('chamfer_config_a', xyz1=x1, xyz2=x2, dist1=d1, idx1=i1, dist2=d2, idx2=i2)
    # the shape of the reference's disabled GPU test (B=32, N=2048 is 1.6 GB of
    # brute force per direction; B=2 keeps the fixture small)
    torch.manual_seed(1)
    x1, x2 = torch.rand(2, 2048, 3), torch.rand(2, 2048, 3)
    d1, i1, d2, i2 = nn_distance_torch(x1, x2, 'NWC')
    npz('chamfer_n2048', seed=1, dist1=d1, idx1=i1, dist2=d2, idx2=i2)

    # 2. SE(3) through the reference's transforms.py
    g = torch.Generator().manual_seed(2)
    q = torch.randn(3, 5, 4, generator=g)
    q[0] = q[0] / q[0].norm(dim=-1, keepdim=True)
    t = torch.randn(3, 5, 3, generator=g)
    v = torch.randn(3, 5, 7, 3, generator=g)
    npz('se3', quat=q, trans=t, pts=v, rot_pc=rot_pc(q, v, rot_type='quat'),
        transform_pc=transform_pc(t, q, v, rot_type='quat'))

    # 3. PointNet (train-mode batch statistics, running-stat update, eval mode)
    enc = fill_params_(build_encoder('pointnet', 256), 3)
    x = torch.randn(6, 50, 3, generator=g)
    enc.train()
    out_train = enc(x)
    rm, rv = enc.bn5.running_mean.clone(), enc.bn5.running_var.clone()
    rm1 = enc.bn1.running_mean.clone()
    enc.eval()
    out_eval = enc(x)
    enc_pp = fill_params_(build_encoder('pointnet', 64, global_feat=False), 4).eval()
    npz('pointnet', x=x, out_train=out_train, bn5_running_mean=rm, bn5_running_var=rv,
        bn1_running_mean=rm1, out_eval_after_update=out_eval, out_perpoint_eval=enc_pp(x))

    # 4. DGCNN
    enc = fill_params_(build_encoder('dgcnn', 128), 5)
    x = torch.rand(3, 64, 3, generator=g) - 0.5
    idx = knn(x.transpose(2, 1).contiguous(), 20).sort(-1)[0]
    enc.train()
    out_train = enc(x)
    enc.eval()
    npz('dgcnn', x=x, knn_sorted=idx, out_train=out_train, out_eval_after_update=enc(x))

    # 5. Transformer encoder (eval: dropout off), with padding
    tr = fill_params_(TransformerEncoder(64, 4, 128, 2), 6).eval()
    tok = torch.randn(3, 6, 64, generator=g)
    mask = torch.tensor([[1, 1, 1, 1, 1, 1], [1, 1, 1, 0, 0, 0], [1, 0, 0, 0, 0, 0]]).bool()
    npz('transformer', tokens=tok, valid=mask, out=tr(tok, mask))
    tr2 = fill_params_(TransformerEncoder(256, 8, 1024, 4), 7).eval()
    tok2 = torch.randn(2, 20, 256, generator=g)
    mask2 = torch.zeros(2, 20).bool()
    mask2[0, :20] = True
    mask2[1, :7] = True
    npz('transformer_full', tokens=tok2, valid=mask2, out=tr2(tok2, mask2))

    # 6. pose head (no noise)
    head = fill_params_(StocasticPoseRegressor(256, 0), 8)
    f = torch.randn(2, 5, 256, generator=g)
    rot, trans = head(f)
    npz('regressor', feats=f, rot=rot, trans=trans)

    # 7. losses and metrics
    batch = make_batch(3, P=6, N=80, num_valid=[6, 3, 1], seed=9)
    pts, valids = batch['part_pcs'], batch['part_valids']
    gq = Rotation3D(batch['part_quat'])
    gt = batch['part_trans']
    pq_raw = torch.randn(3, 6, 4, generator=g)
    pq = Rotation3D(pq_raw / pq_raw.norm(dim=-1, keepdim=True))
    pt = torch.randn(3, 6, 3, generator=g) * 0.2
    npz('losses', pred_quat=pq.rot, pred_trans=pt,
        trans_l2=trans_l2_loss(pt, gt, valids), rot_l2=rot_l2_loss(pq, gq, valids),
        rot_cosine=rot_cosine_loss(pq, gq, valids),
        rot_points_l2=rot_points_l2_loss(pts, pq, gq, valids),
        rot_points_cd=rot_points_cd_loss(pts, pq, gq, valids),
        shape_cd_train=shape_cd_loss(pts, pt, gt, pq, gq, valids, training=True),
        shape_cd_eval=shape_cd_loss(pts, pt, gt, pq, gq, valids, training=False),
        part_acc=calc_part_acc(pts, pt, gt, pq, gq, valids),
        part_acc_close=calc_part_acc(pts, gt + 0.01, gt, gq, gq, valids),
        trans_rmse=trans_metrics(pt, gt, valids, 'rmse'), rot_mae=rot_metrics(pq, gq, valids, 'mae'),
        rot_rmse=rot_metrics(pq, gq, valids, 'rmse'))

    # 8. whole models: loss dicts of training_step (dropout 0) and validation_step
    def run_model(cfg_rel, tag, seed, semantic=False, encoder=None, B=2, N=64, nv=(5, 3)):
        cfg = load_ref_cfg(cfg_rel)
        if encoder:
            cfg.model.encoder = encoder
        model = zero_dropout(fill_params_(build_model(cfg), seed))
        model.trainer = lightning.Trainer()
        out = {}
        for mode in ('train', 'val'):
            batch = make_batch(B, P=20, N=N, num_valid=list(nv), seed=seed, semantic=semantic)
            if semantic:
                cp = torch.zeros(B, 20, 20, 4)
                cp[:, 0, 1, 0] = cp[:, 1, 0, 0] = 1
                cp[:, 0, 1, 1:] = 0.1
                cp[:, 1, 0, 1:] = -0.1
                batch['contact_points'] = cp
            model.train(mode == 'train')
            torch.manual_seed(100 + seed)
            with torch.set_grad_enabled(mode == 'train'):
                ld = model.forward_pass(batch, mode=mode, optimizer_idx=-1)
            for k, v in ld.items():
                out[f'{mode}/{k}'] = torch.as_tensor(v).float()
            if mode == 'train':
                ld['loss'].backward()
                gn = torch.sqrt(sum((p.grad.double()**2).sum() for p in model.parameters()
                                    if p.grad is not None))
                out['train/grad_norm'] = gn.float()
                model.zero_grad()
        npz(tag, **out)

    run_model('pn_transformer/pn_transformer/pn_transformer-32x1-cosine_400e-everyday.py',
              'model_pn_transformer', 11)
    run_model('pn_transformer/pn_transformer/pn_transformer-32x1-cosine_400e-partnet_chair.py',
              'model_pn_transformer_semantic', 12, semantic=True, nv=(7, 4), N=128)
    run_model('global/global-32x1-cosine_200e-everyday.py', 'model_global', 13)
    run_model('dgl/dgl-32x1-cosine_200e-everyday.py', 'model_dgl', 14)
    run_model('dgl/dgl-32x1-cosine_200e-everyday.py', 'model_dgl_dgcnn', 15, encoder='dgcnn')
    run_model('pn_transformer/pn_transformer_refine/pn_transformer_refine-32x1-cosine_400e-everyday.py',
              'model_pn_transformer_refine', 16)


if __name__ == '__main__':
    main()
