"""The CPU oracle against the golden vectors produced from the reference
itself (oracle/make_golden.py) -- this is what pins the oracle (prompt rule 3).
No GPU needed."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import cpu, torch_ref
from oracle.params import fill_params_

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def test_chamfer_config_a_vs_reference_definition():
    g = gold('chamfer_config_a')
    for fused in (True, False):
        d1, i1, d2, i2 = cpu.chamfer_forward(g['xyz1'], g['xyz2'], fused=fused)
        # tolerance and exact indices of the reference's own test (test_chamfer.py:72-76)
        np.testing.assert_allclose(d1, g['dist1'], atol=1e-6)
        np.testing.assert_allclose(d2, g['dist2'], atol=1e-6)
        np.testing.assert_array_equal(i1, g['idx1'])
        np.testing.assert_array_equal(i2, g['idx2'])
    # the un-contracted restatement is bit-identical to the torch definition
    d1, _, d2, _ = cpu.chamfer_forward(g['xyz1'], g['xyz2'], fused=False)
    assert np.max(np.abs(d1 - g['dist1'])) < 1e-7


def test_chamfer_reference_test_shape():
    g = gold('chamfer_n2048')
    torch.manual_seed(int(g['seed']))
    x1, x2 = torch.rand(2, 2048, 3).numpy(), torch.rand(2, 2048, 3).numpy()
    d1, i1, d2, i2 = cpu.chamfer_forward(x1, x2)
    np.testing.assert_allclose(d1, g['dist1'], atol=1e-6)
    np.testing.assert_array_equal(i1, g['idx1'])
    np.testing.assert_array_equal(i2, g['idx2'])


def test_chamfer_backward_vs_float64_definition():
    rng = np.random.default_rng(0)
    x1 = rng.random((2, 64, 3)).astype(np.float32)
    x2 = rng.random((2, 64, 3)).astype(np.float32)
    a = torch.from_numpy(x1).double().requires_grad_()
    b = torch.from_numpy(x2).double().requires_grad_()
    D = ((a[:, :, None] - b[:, None])**2).sum(-1)
    w1 = torch.from_numpy(rng.random((2, 64))).double()
    w2 = torch.from_numpy(rng.random((2, 64))).double()
    ((D.min(2)[0] * w1).sum() + (D.min(1)[0] * w2).sum()).backward()
    _, i1, _, i2 = cpu.chamfer_forward(x1, x2)
    g1, g2 = cpu.chamfer_backward(w1.float().numpy(), w2.float().numpy(), x1, x2, i1, i2)
    np.testing.assert_allclose(g1, a.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(g2, b.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_se3_vs_reference_transforms():
    g = gold('se3')
    np.testing.assert_array_equal(cpu.se3_transform(g['quat'], None, g['pts']), g['rot_pc'])
    np.testing.assert_array_equal(cpu.se3_transform(g['quat'], g['trans'], g['pts']),
                                  g['transform_pc'])
    out = torch_ref.qtransform(*[torch.from_numpy(g[k]) for k in ('trans', 'quat', 'pts')])
    np.testing.assert_array_equal(out.numpy(), g['transform_pc'])


def test_se3_vs_scipy():
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(1)
    q = rng.standard_normal((5, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    v = rng.standard_normal((5, 9, 3)).astype(np.float32)
    want = np.stack([R.from_quat(q[i, [1, 2, 3, 0]]).apply(v[i]) for i in range(5)])
    np.testing.assert_allclose(cpu.se3_transform(q, None, v), want, rtol=1e-5, atol=1e-6)


def _sd(module):
    return {k: v.detach() for k, v in module.state_dict().items()}


def test_pointnet_restatement():
    from multi_part_assembly_b200.models import build_encoder
    g = gold('pointnet')
    x = torch.from_numpy(g['x'])
    sd = _sd(fill_params_(build_encoder('pointnet', 256), 3))
    out = torch_ref.pointnet_forward(x, sd, training=True)
    np.testing.assert_allclose(out.numpy(), g['out_train'], rtol=1e-4, atol=1e-5)
    sd = _sd(fill_params_(build_encoder('pointnet', 64, global_feat=False), 4))
    out = torch_ref.pointnet_forward(x, sd, training=False, global_feat=False)
    np.testing.assert_allclose(out.numpy(), g['out_perpoint_eval'], rtol=1e-4, atol=1e-5)


def test_knn_restatement_sets():
    g = gold('dgcnn')
    x = np.ascontiguousarray(g['x'].transpose(0, 2, 1))
    np.testing.assert_array_equal(cpu.knn(x, 20), g['knn_sorted'])


def test_dgcnn_restatement():
    from multi_part_assembly_b200.models import build_encoder
    g = gold('dgcnn')
    sd = _sd(fill_params_(build_encoder('dgcnn', 128), 5))
    out = torch_ref.dgcnn_forward(torch.from_numpy(g['x']), sd, training=True)
    np.testing.assert_allclose(out.numpy(), g['out_train'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('name,dims,seed', [('transformer', (64, 4, 128, 2), 6),
                                            ('transformer_full', (256, 8, 1024, 4), 7)])
def test_transformer_restatement(name, dims, seed):
    from multi_part_assembly_b200.models.pn_transformer import TransformerEncoder
    g = gold(name)
    sd = _sd(fill_params_(TransformerEncoder(*dims), seed))
    out = torch_ref.transformer_forward(torch.from_numpy(g['tokens']), torch.from_numpy(g['valid']),
                                        sd, dims[1], dims[3])
    valid = g['valid']
    np.testing.assert_allclose(out.numpy()[valid], g['out'][valid], rtol=1e-4, atol=1e-5)


def test_regressor_restatement():
    from multi_part_assembly_b200.models import StocasticPoseRegressor
    g = gold('regressor')
    sd = _sd(fill_params_(StocasticPoseRegressor(256, 0), 8))
    rot, trans = torch_ref.pose_regressor_forward(torch.from_numpy(g['feats']), sd)
    np.testing.assert_allclose(rot.numpy(), g['rot'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(trans.numpy(), g['trans'], rtol=1e-5, atol=1e-6)


def test_losses_restatement():
    from multi_part_assembly_b200.datasets import make_batch
    g = gold('losses')
    b = make_batch(3, P=6, N=80, num_valid=[6, 3, 1], seed=9)
    pts, valids, gt = b['part_pcs'], b['part_valids'], b['part_trans']
    gq = torch_ref.process_zero_quat(b['part_quat'])
    pq, pt = torch.from_numpy(g['pred_quat']), torch.from_numpy(g['pred_trans'])
    got = {
        'trans_l2': torch_ref.trans_l2_loss(pt, gt, valids),
        'rot_cosine': torch_ref.rot_cosine_loss(pq, gq, valids),
        'rot_points_l2': torch_ref.rot_points_l2_loss(pts, pq, gq, valids),
        'rot_points_cd': torch_ref.rot_points_cd_loss(pts, pq, gq, valids),
        'shape_cd_train': torch_ref.shape_cd_loss(pts, pt, gt, pq, gq, valids, True),
        'shape_cd_eval': torch_ref.shape_cd_loss(pts, pt, gt, pq, gq, valids, False),
        'part_acc': torch_ref.calc_part_acc(pts, pt, gt, pq, gq, valids),
    }
    for k, v in got.items():
        np.testing.assert_allclose(v.numpy(), g[k], rtol=1e-6, atol=1e-7, err_msg=k)


def test_full_pn_transformer_restatement():
    """oracle pn_transformer forward + geometric losses == the reference
    model's training_step loss dict (dropout 0)."""
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    g = gold('model_pn_transformer')
    sd = _sd(fill_params_(build_model(get_cfg('pn_transformer')), 11))
    batch = make_batch(2, P=20, N=64, num_valid=[5, 3], seed=11)
    rot, trans = torch_ref.pn_transformer_forward(batch, sd, training=True)
    out, _ = torch_ref.geometric_losses(batch, rot, trans, training=True)
    for k, v in out.items():
        np.testing.assert_allclose(v.detach().numpy(), g[f'train/{k}'], rtol=2e-4, atol=1e-6,
                                   err_msg=k)


def test_c_abi_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly the entry
    points include/mpa_b200.h declares (no compute calls here)."""
    import ctypes
    from multi_part_assembly_b200 import _lib
    header = open(os.path.join(os.path.dirname(GOLD), '..', 'include', 'mpa_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(mpa_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 10
    handle = ctypes.CDLL(_lib.library_path())
    for name in declared:
        assert hasattr(handle, name), name
    assert declared == set(_lib._SIGNATURES), declared ^ set(_lib._SIGNATURES)
    assert _lib.lib().mpa_version() >= 100


def test_lsap_restatement_matches_scipy():
    """oracle/lsap.py (the algorithm the CUDA assignment kernel follows) is pinned against the
    installed SciPy's linear_sum_assignment -- the call the reference makes
    (base_model.py:175-176) -- on random, tie-heavy and constant cost matrices."""
    from scipy.optimize import linear_sum_assignment
    from oracle.lsap import linear_sum_assignment_square
    rng = np.random.default_rng(0)
    for trial in range(600):
        n = int(rng.integers(1, 21))
        kind = trial % 4
        if kind == 0:
            c = rng.random((n, n)).astype(np.float32)
        elif kind == 1:
            c = rng.integers(0, 4, (n, n)).astype(np.float32)
        elif kind == 2:
            c = np.full((n, n), 0.5, np.float32)
        else:
            c = (rng.random((n, n)) * 1e-3 + rng.integers(0, 2, (n, n))).astype(np.float32)
        want = linear_sum_assignment(c)[1]
        got = linear_sum_assignment_square(c.astype(np.float64))
        assert list(want) == list(got), (trial, n)
