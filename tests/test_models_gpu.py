"""Product modules, losses and whole models on the GPU against the golden
vectors generated from the reference (tests/golden, oracle/make_golden.py).
Tolerance: north_star asks 1e-5 rel for Chamfer/pose losses given the same
poses; network outputs accumulate in a different order than the CPU
reference, so module outputs use 1e-4 rel."""
import os

import numpy as np
import pytest
import torch

from oracle.params import fill_params_, zero_dropout

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def T(a, dev):
    return torch.from_numpy(np.asarray(a)).to(dev)


def test_pointnet_module(cuda):
    from multi_part_assembly_b200.models import build_encoder
    g = gold('pointnet')
    enc = fill_params_(build_encoder('pointnet', 256), 3).to(cuda).train()
    x = T(g['x'], cuda)
    out = enc(x)
    np.testing.assert_allclose(out.cpu().detach().numpy(), g['out_train'], rtol=1e-4, atol=1e-5)
    # train-mode side effect: running statistics (momentum 0.1, unbiased var)
    np.testing.assert_allclose(enc.bn5.running_mean.cpu().numpy(), g['bn5_running_mean'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(enc.bn5.running_var.cpu().numpy(), g['bn5_running_var'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(enc.bn1.running_mean.cpu().numpy(), g['bn1_running_mean'], rtol=1e-4, atol=1e-6)
    enc.eval()
    np.testing.assert_allclose(enc(x).cpu().detach().numpy(), g['out_eval_after_update'], rtol=1e-4, atol=1e-5)
    pp = fill_params_(build_encoder('pointnet', 64, global_feat=False), 4).to(cuda).eval()
    np.testing.assert_allclose(pp(x).cpu().detach().numpy(), g['out_perpoint_eval'], rtol=1e-4, atol=1e-5)


def test_dgcnn_module(cuda):
    from multi_part_assembly_b200.models import build_encoder
    g = gold('dgcnn')
    enc = fill_params_(build_encoder('dgcnn', 128), 5).to(cuda).train()
    x = T(g['x'], cuda)
    np.testing.assert_allclose(enc(x).cpu().detach().numpy(), g['out_train'], rtol=1e-4, atol=1e-5)
    enc.eval()
    np.testing.assert_allclose(enc(x).cpu().detach().numpy(), g['out_eval_after_update'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('name,dims,seed', [('transformer', (64, 4, 128, 2), 6),
                                            ('transformer_full', (256, 8, 1024, 4), 7)])
def test_transformer_module(cuda, name, dims, seed):
    from multi_part_assembly_b200.models.pn_transformer import TransformerEncoder
    g = gold(name)
    tr = fill_params_(TransformerEncoder(*dims), seed).to(cuda).eval()
    out = tr(T(g['tokens'], cuda), T(g['valid'], cuda)).cpu().detach().numpy()
    v = g['valid']
    # fp32 mode = the fp32-accurate tensor-core GEMMs (three bf16 planes per operand): ~3e-6 of
    # the output scale per GEMM (the tensor core's fp32 accumulator truncates), 2e-5 after 4 layers
    np.testing.assert_allclose(out[v], g['out'][v], rtol=1e-4, atol=4e-5)
    assert np.isfinite(out).all()  # padded slots must stay finite


def test_losses(cuda):
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200 import utils as U
    g = gold('losses')
    b = make_batch(3, P=6, N=80, num_valid=[6, 3, 1], seed=9, device=cuda)
    pts, valids, gt = b['part_pcs'], b['part_valids'], b['part_trans']
    gq = U.Rotation3D(b['part_quat'])
    pq, pt = U.Rotation3D(T(g['pred_quat'], cuda)), T(g['pred_trans'], cuda)
    got = {
        'trans_l2': U.trans_l2_loss(pt, gt, valids), 'rot_l2': U.rot_l2_loss(pq, gq, valids),
        'rot_cosine': U.rot_cosine_loss(pq, gq, valids),
        'rot_points_l2': U.rot_points_l2_loss(pts, pq, gq, valids),
        'rot_points_cd': U.rot_points_cd_loss(pts, pq, gq, valids),
        'shape_cd_train': U.shape_cd_loss(pts, pt, gt, pq, gq, valids, training=True),
        'shape_cd_eval': U.shape_cd_loss(pts, pt, gt, pq, gq, valids, training=False),
        'part_acc': U.calc_part_acc(pts, pt, gt, pq, gq, valids),
        'part_acc_close': U.calc_part_acc(pts, gt + 0.01, gt, gq, gq, valids),
        'trans_rmse': U.trans_metrics(pt, gt, valids, 'rmse'),
        'rot_mae': U.rot_metrics(pq, gq, valids, 'mae'),
        'rot_rmse': U.rot_metrics(pq, gq, valids, 'rmse'),
    }
    for k, v in got.items():
        tol = 1e-4 if k.startswith('rot_m') or k.startswith('rot_r') else 1e-5
        np.testing.assert_allclose(v.cpu().numpy(), g[k], rtol=tol, atol=1e-7, err_msg=k)


CASES = [
    ('model_pn_transformer', 'pn_transformer', 'everyday', None, 11, False, 64, (5, 3)),
    ('model_pn_transformer_semantic', 'pn_transformer', 'partnet_chair', None, 12, True, 128, (7, 4)),
    ('model_global', 'global', 'everyday', None, 13, False, 64, (5, 3)),
    ('model_dgl', 'dgl', 'everyday', None, 14, False, 64, (5, 3)),
    ('model_dgl_dgcnn', 'dgl', 'everyday', 'dgcnn', 15, False, 64, (5, 3)),
    ('model_pn_transformer_refine', 'pn_transformer_refine', 'everyday', None, 16, False, 64, (5, 3)),
]


@pytest.mark.parametrize('tag,name,dataset,encoder,seed,semantic,N,nv', CASES)
def test_model_steps_match_reference(cuda, tag, name, dataset, encoder, seed, semantic, N, nv):
    """training_step / validation_step loss dicts and the gradient norm of the
    reference model (same weights, same batch, same RNG seed, dropout 0)."""
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    from multi_part_assembly_b200.compat.lightning import Trainer
    g = gold(tag)
    cfg = get_cfg(name, dataset)
    if encoder:
        cfg.model.encoder = encoder
    model = zero_dropout(fill_params_(build_model(cfg), seed)).to(cuda)
    model.trainer = Trainer()
    for mode in ('train', 'val'):
        batch = make_batch(2, P=20, N=N, num_valid=list(nv), seed=seed, semantic=semantic)
        if semantic:
            cp = torch.zeros(2, 20, 20, 4)
            cp[:, 0, 1, 0] = cp[:, 1, 0, 0] = 1
            cp[:, 0, 1, 1:] = 0.1
            cp[:, 1, 0, 1:] = -0.1
            batch['contact_points'] = cp
        batch = {k: v.to(cuda) for k, v in batch.items()}
        model.train(mode == 'train')
        torch.manual_seed(100 + seed)
        with torch.set_grad_enabled(mode == 'train'):
            ld = model.forward_pass(batch, mode=mode, optimizer_idx=-1)
        for k, v in ld.items():
            want = g[f'{mode}/{k}']
            got = float(v)
            assert np.isfinite(got), k
            np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-6, err_msg=f'{mode}/{k}')
        if mode == 'train':
            ld['loss'].backward()
            gn = torch.sqrt(sum((p.grad.double()**2).sum() for p in model.parameters()
                                if p.grad is not None))
            np.testing.assert_allclose(float(gn), g['train/grad_norm'], rtol=2e-3)
            model.zero_grad()


def _bf16_emulated_pointnet(x, enc, training):
    """The arithmetic of csrc/pointnet.cu in plain torch: bf16-rounded weights
    and layer inputs, fp32 accumulation, fp32 BatchNorm on the fp32 conv output."""
    import torch.nn.functional as F
    q = lambda t: t.to(torch.bfloat16).float()
    h = q(x.transpose(2, 1))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for i in range(1, 6):
            conv, bn = getattr(enc, f'conv{i}'), getattr(enc, f'bn{i}')
            y = F.conv1d(h, q(conv.weight))
            y = F.batch_norm(y, None if training else bn.running_mean,
                             None if training else bn.running_var, bn.weight, bn.bias, training, 0.0, bn.eps)
            h = q(F.relu(y)) if i < 5 else y
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    return h.max(dim=-1)[0]


@pytest.mark.parametrize('feat,n,N', [(256, 6, 50), (256, 40, 1000), (128, 9, 333), (128, 3, 128)])
@pytest.mark.parametrize('training', [True, False])
def test_pointnet_tcgen05_kernel(cuda, feat, n, N, training):
    """Native tcgen05 PointNet vs (a) the same bf16 arithmetic in torch (tight)
    and (b) the fp32 module (loose: bf16 operand rounding)."""
    import copy
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models import build_encoder
    enc = fill_params_(build_encoder('pointnet', feat), 21).to(cuda).train(training)
    ref = copy.deepcopy(enc)
    g = torch.Generator().manual_seed(n * N)
    x = (torch.rand(n, N, 3, generator=g) - 0.5).to(cuda)
    valids = torch.ones(n, device=cuda)
    valids[n // 2] = 0
    kernels.set_precision('bf16')
    try:
        out = enc(x, valids=valids)
    finally:
        kernels.set_precision('auto')
    keep = valids.bool()
    want = _bf16_emulated_pointnet(x[keep], ref, training)
    np.testing.assert_allclose(out[keep].detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=2e-2, atol=2e-2)
    assert torch.all(out[~keep] == 0)
    kernels.set_precision('fp32')
    try:
        full = ref(x[keep])
    finally:
        kernels.set_precision('auto')
    err = (out[keep] - full).abs().max().item() / full.abs().max().item()
    assert err < 5e-2, err
    if training:  # running statistics updated like nn.BatchNorm1d (momentum 0.1, unbiased var)
        np.testing.assert_allclose(enc.bn1.running_mean.cpu().numpy(), ref.bn1.running_mean.cpu().numpy(),
                                   rtol=2e-2, atol=2e-3)
        np.testing.assert_allclose(enc.bn5.running_var.cpu().numpy(), ref.bn5.running_var.cpu().numpy(),
                                   rtol=5e-2, atol=5e-3)
        assert int(enc.bn3.num_batches_tracked) == 1


@pytest.mark.parametrize('M,N,K,act,res', [(640, 768, 256, 0, False), (640, 256, 1024, 0, True),
                                           (40, 1024, 256, 1, False), (7, 4, 128, 2, False),
                                           (129, 136, 288, 2, True), (5120, 256, 256, 0, True)])
def test_linear_tcgen05(cuda, M, N, K, act, res):
    """TMA-fed tcgen05 GEMM vs the same bf16-operand / fp32-accumulate product in torch."""
    from multi_part_assembly_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K**0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    r = torch.randn(M, N, generator=g).to(cuda) if res else None
    out = torch.empty(M, N, device=cuda)
    L = _lib.lib()
    rc = L.mpa_linear_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(r), M, N, K, act,
                              _lib.ptr(out), None, 0, _lib.cuda_stream(cuda))
    _lib.check(rc, 'mpa_linear_forward')
    want = x.to(torch.bfloat16).double() @ w.to(torch.bfloat16).double().T + b.double()
    if act == 1:
        want = torch.relu(want)
    elif act == 2:
        want = torch.nn.functional.leaky_relu(want, 0.2)
    if res:
        want = want + r.double()
    np.testing.assert_allclose(out.cpu().numpy(), want.float().cpu().numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('name,dims,seed', [('transformer', (64, 4, 128, 2), 6),
                                            ('transformer_full', (256, 8, 1024, 4), 7)])
def test_transformer_tcgen05(cuda, name, dims, seed):
    """Native transformer encoder (bf16 tensor-core GEMMs) vs the reference golden output."""
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models.pn_transformer import TransformerEncoder
    g = gold(name)
    tr = fill_params_(TransformerEncoder(*dims), seed).to(cuda).eval()
    kernels.set_precision('bf16')
    try:
        out = tr(T(g['tokens'], cuda), T(g['valid'], cuda)).cpu().detach().numpy()
    finally:
        kernels.set_precision('auto')
    v = g['valid']
    assert np.isfinite(out).all()
    err = np.abs(out[v] - g['out'][v]).max() / np.abs(g['out'][v]).max()
    assert err < 3e-2, err


@pytest.mark.parametrize('tag,name,dataset,seed,semantic,N,nv', [
    ('model_pn_transformer', 'pn_transformer', 'everyday', 11, False, 64, (5, 3)),
    ('model_global', 'global', 'everyday', 13, False, 64, (5, 3)),
    # models that score several predictions per step (GNN / refine iterations): the totals
    # are sums over the iterations, which the single-prediction shortcut must not replace
    ('model_dgl', 'dgl', 'everyday', 14, False, 64, (5, 3)),
    ('model_pn_transformer_refine', 'pn_transformer_refine', 'everyday', 16, False, 64, (5, 3))])
def test_fused_loss_path_matches_reference(cuda, tag, name, dataset, seed, semantic, N, nv):
    """forward_pass without autograd (the benchmark / CUDA-graph path) uses the
    fused loss kernels; its loss dict must equal the reference's too."""
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    from multi_part_assembly_b200.compat.lightning import Trainer
    g = gold(tag)
    model = zero_dropout(fill_params_(build_model(get_cfg(name, dataset)), seed)).to(cuda)
    model.trainer = Trainer()
    for mode in ('train', 'val'):
        batch = {k: v.to(cuda) for k, v in
                 make_batch(2, P=20, N=N, num_valid=list(nv), seed=seed, semantic=semantic).items()}
        model.train(mode == 'train')
        torch.manual_seed(100 + seed)
        with torch.no_grad():
            ld = model.forward_pass(batch, mode=mode, optimizer_idx=-1)
        for k, v in ld.items():
            np.testing.assert_allclose(float(v), g[f'{mode}/{k}'], rtol=2e-4, atol=1e-6,
                                       err_msg=f'{mode}/{k}')


def test_pose_head_native(cuda):
    """Native pose head (one fused fp32 kernel) vs the golden reference output."""
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models import StocasticPoseRegressor
    g = gold('regressor')
    head = fill_params_(StocasticPoseRegressor(256, 0), 8).to(cuda)
    kernels.set_precision('bf16')
    try:
        with torch.no_grad():
            rot, trans = head(T(g['feats'], cuda))
    finally:
        kernels.set_precision('auto')
    np.testing.assert_allclose(rot.cpu().numpy(), g['rot'], rtol=1e-4, atol=1e-5)  # fp32 kernel
    np.testing.assert_allclose(trans.cpu().numpy(), g['trans'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rot.norm(dim=-1).cpu().numpy(), 1.0, rtol=1e-5)


@pytest.mark.parametrize('feat,T', [(276, 37), (256, 640), (512, 9), (64, 1)])
def test_pose_head_kernel_shapes(cuda, feat, T):
    """Fused pose head on ragged token counts / input widths vs the stock fp32 layers."""
    from multi_part_assembly_b200.models import StocasticPoseRegressor
    head = fill_params_(StocasticPoseRegressor(feat, 0), 31).to(cuda)
    x = torch.randn(T, feat, generator=torch.Generator().manual_seed(T)).to(cuda)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.enable_grad():
            want_rot, want_trans = head(x)  # autograd path: stock torch layers
        with torch.no_grad():
            rot, trans = head(x)            # native kernel
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    np.testing.assert_allclose(rot.cpu().numpy(), want_rot.detach().cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(trans.cpu().numpy(), want_trans.detach().cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('feat,n,N,with_valids', [(256, 6, 100, False), (128, 9, 333, True), (256, 40, 1000, True)])
def test_pointnet_native_backward(cuda, feat, n, N, with_valids):
    """Hand-written PointNet backward (streaming BatchNorm/ReLU/max-pool kernels + library
    GEMMs, bf16 activations) vs fp32 autograd through the stock layers.  Bar: within 6e-2
    relative L2 per parameter, or -- on tiny batches, where five train-mode BatchNorms
    amplify bf16 rounding -- no worse than 2x the error of stock bf16-autocast autograd."""
    import copy
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models import build_encoder
    enc = fill_params_(build_encoder('pointnet', feat), 5).to(cuda).train()
    stock = copy.deepcopy(enc)
    ref = copy.deepcopy(enc)
    g = torch.Generator().manual_seed(n + N)
    x = (torch.rand(n, N, 3, generator=g) - 0.5).to(cuda)
    w = torch.randn(n, feat, generator=g).to(cuda)
    valids = None
    if with_valids:
        valids = torch.ones(n, device=cuda)
        valids[1] = 0
        valids[n - 1] = 0

    def run(model, native):
        kernels.set_precision('bf16')
        kernels._NATIVE_BACKWARD['pointnet'] = native
        try:
            out = model(x, valids=valids) if with_valids else model(x)
            (out * w).sum().backward()
        finally:
            kernels.set_precision('auto')
            kernels._NATIVE_BACKWARD['pointnet'] = True

    run(enc, True)
    run(stock, False)
    kernels.set_precision('fp32')
    try:
        keep = valids.bool() if with_valids else torch.ones(n, dtype=torch.bool, device=cuda)
        (ref(x[keep]) * w[keep]).sum().backward()
    finally:
        kernels.set_precision('auto')
    for (name, p), (_, s_), (_, q) in zip(enc.named_parameters(), stock.named_parameters(),
                                          ref.named_parameters()):
        assert p.grad is not None, name
        a, b, c = p.grad.flatten().double(), s_.grad.flatten().double(), q.grad.flatten().double()
        rel = float((a - c).norm() / c.norm().clamp_min(1e-12))
        rel_stock = float((b - c).norm() / c.norm().clamp_min(1e-12))
        assert rel < max(6e-2, 2.0 * rel_stock), (name, rel, rel_stock)


def test_graphed_train_step_matches_eager(cuda):
    """runtime.GraphedTrainStep (forward, loss, backward, Adam in one CUDA graph) follows the
    same optimisation trajectory as the eager loop it replaces."""
    import copy
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    from multi_part_assembly_b200.compat.lightning import Trainer
    from multi_part_assembly_b200.runtime import GraphedTrainStep
    model = zero_dropout(fill_params_(build_model(get_cfg('pn_transformer', 'everyday')), 3)).to(cuda).train()
    model.trainer = Trainer()
    eager = copy.deepcopy(model)
    eager.trainer = Trainer()
    batch = make_batch(4, P=20, N=128, num_valid=[5, 20, 3, 9], seed=1, device=cuda)

    opt_e = eager.configure_optimizers()
    opt_e = opt_e[0][0] if isinstance(opt_e, tuple) else opt_e
    want = []
    for _ in range(5):  # 3 warm-up steps inside GraphedTrainStep + 2 replays
        with torch.autocast('cuda', dtype=torch.bfloat16):
            loss = eager.training_step(dict(batch), 0)
        opt_e.zero_grad(set_to_none=True)
        loss.backward()
        opt_e.step()
        want.append(float(loss))

    opt_g = model.configure_optimizers()
    opt_g = opt_g[0][0] if isinstance(opt_g, tuple) else opt_g
    g = GraphedTrainStep(model, opt_g, batch, warmup=3)  # 3 eager steps, then the capture (records only)
    got = [float(g()) for _ in range(2)]
    np.testing.assert_allclose(got, want[3:5], rtol=5e-3)
    assert all(np.isfinite(got))


def test_transformer_native_dropout(cuda):
    """Training with the reference's dropout 0.1 (transformer.py:10,47) stays on the native
    kernels: the Philox masks drawn in-kernel are returned, so the forward can be checked
    against the same layer chain in fp32 torch WITH THOSE MASKS (bf16 bar), the keep rate
    against 1 - p, reseeding against reproducibility, and the backward against fp32 autograd."""
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models.pn_transformer import TransformerEncoder
    B, P, D, H, FF, Lyr, p = 6, 20, 256, 8, 1024, 4, 0.1
    tr = fill_params_(TransformerEncoder(D, H, FF, Lyr), 7).to(cuda).train()
    g = torch.Generator().manual_seed(3)
    tokens = torch.randn(B, P, D, generator=g).to(cuda).requires_grad_(True)
    valid = torch.ones(B, P, dtype=torch.bool, device=cuda)
    valid[1, 7:] = False
    valid[4, 2:] = False

    def run(seed):
        torch.manual_seed(seed)
        kernels._TRANSFORMER_TRACE = trace = []
        kernels.set_precision('bf16')
        try:
            out = tr(tokens, valid)
        finally:
            kernels.set_precision('auto')
            kernels._TRANSFORMER_TRACE = None
        assert len(trace) == 1 and trace[0] is not None  # the native path ran and drew masks
        return out, trace[0]

    out, masks = run(5)
    ms = kernels.split_transformer_masks(masks, B, P, D, H, FF, Lyr)
    for layer in ms:
        for m in layer:
            assert abs(float(m.float().mean()) - (1 - p)) < 0.02
    enc = tr.transformer_encoder
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        tok32 = tokens.detach().clone().requires_grad_(True)
        want = kernels._transformer_masked_torch(tok32, valid, enc, H, ms, p)
        w = torch.randn(B, P, D, generator=g).to(cuda) * valid[..., None]
        (want * w).sum().backward()
        want_grads = [tok32.grad] + [q.grad.clone() for q in enc.parameters()]
        enc.zero_grad()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    v = valid.cpu().numpy()
    a, b = out.detach().cpu().numpy(), want.detach().cpu().numpy()
    assert np.isfinite(a).all()
    assert np.abs(a[v] - b[v]).max() / np.abs(b[v]).max() < 3e-2
    # backward through the saved masks vs fp32 autograd with the same masks
    (out * w).sum().backward()
    got_grads = [tokens.grad] + [q.grad for q in enc.parameters()]
    for gg, ww in zip(got_grads, want_grads):
        # the backward runs the layer chain under bf16 autocast (like the stock path it
        # replaces): bf16 rounding of activations and gradients, ~7 % on the smallest tensors
        rel = float((gg.double() - ww.double()).norm() / ww.double().norm().clamp_min(1e-12))
        assert rel < 1e-1, rel
    # a second forward draws new masks; re-seeding reproduces the first ones
    _, masks2 = run(5)
    assert torch.equal(masks, masks2)
    kernels._TRANSFORMER_TRACE = trace = []
    kernels.set_precision('bf16')
    try:
        tr(tokens, valid)
    finally:
        kernels.set_precision('auto')
        kernels._TRANSFORMER_TRACE = None
    assert not torch.equal(trace[0], masks)
    # eval mode: no dropout, no masks
    tr.eval()
    kernels._TRANSFORMER_TRACE = trace = []
    kernels.set_precision('bf16')
    try:
        tr(tokens, valid)
    finally:
        kernels.set_precision('auto')
        kernels._TRANSFORMER_TRACE = None
    assert trace == [None]


@pytest.mark.parametrize('M,N,K,act,res', [(640, 768, 256, 0, False), (640, 256, 1024, 0, True),
                                           (129, 136, 288, 2, True), (5000, 128, 6, 0, False),
                                           (1000, 512, 512, 1, False)])
def test_linear_fp32_accurate_mode(cuda, M, N, K, act, res):
    """MPA_PRECISION_FP32: three bf16 planes per operand, six tensor-core products per k-step
    -- agrees with a float64 product of the fp32 operands like an fp32 GEMM does (bar: 8x the
    error of torch's own fp32 matmul, floor 5e-6 of the output scale; measured ~3e-6: the
    tensor core's fp32 accumulator truncates)."""
    from multi_part_assembly_b200 import kernels
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K**0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    r = torch.randn(M, N, generator=g).to(cuda) if res else None
    out = kernels.linear(x, w, b, act=act, residual=r, precision=kernels.PRECISION_FP32)
    want = x.double() @ w.double().T + b.double()
    ref32 = x @ w.T + b
    if act == 1:
        want, ref32 = torch.relu(want), torch.relu(ref32)
    elif act == 2:
        want = torch.nn.functional.leaky_relu(want, 0.2)
        ref32 = torch.nn.functional.leaky_relu(ref32, 0.2)
    if res:
        want, ref32 = want + r.double(), ref32 + r
    scale = float(want.abs().max())
    err = float((out.double() - want).abs().max())
    err32 = float((ref32.double() - want).abs().max())
    assert err <= max(8 * err32, 5e-6 * scale), (err, err32, scale)


@pytest.mark.parametrize('B', [3, 100, 520])
def test_transformer_cluster_widths_agree(cuda, B):
    """The fused FFN block splits the hidden dimension over a thread-block cluster when there are
    few token tiles (B = 3: 8 CTAs per tile, B = 100: 8, B = 520: one CTA per tile).  All widths
    compute the same function: compared with the fp32 torch layer chain at the bf16 bar, and
    the batch-size independence of a shape's output is checked across widths."""
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models.pn_transformer import TransformerEncoder
    P, D, H, FF, Lyr = 20, 256, 8, 1024, 4
    tr = fill_params_(TransformerEncoder(D, H, FF, Lyr), 11).to(cuda).eval()
    g = torch.Generator().manual_seed(B)
    tokens = torch.randn(B, P, D, generator=g).to(cuda)
    valid = torch.ones(B, P, dtype=torch.bool, device=cuda)
    valid[0, 13:] = False
    kernels.set_precision('bf16')
    try:
        with torch.no_grad():
            out = tr(tokens, valid)
            small = tr(tokens[:3].contiguous(), valid[:3].contiguous())  # 8-wide clusters
    finally:
        kernels.set_precision('auto')
    enc = tr.transformer_encoder if hasattr(tr, 'transformer_encoder') else tr
    enc = next(m for m in tr.modules() if isinstance(m, torch.nn.TransformerEncoder))
    with torch.no_grad():
        want = kernels._transformer_masked_torch(tokens, valid, enc, H, None, 0.0)
    v = valid.cpu().numpy()
    a, b = out.cpu().numpy(), want.cpu().numpy()
    assert np.isfinite(a).all()
    assert np.abs(a[v] - b[v]).max() / np.abs(b[v]).max() < 3e-2
    # a shape's tokens only see that shape: the first three shapes do not depend on the batch size.
    # Different cluster widths add the hidden-dimension partial sums in a different order; a last-bit
    # difference can flip the bf16 rounding of the next layer's operand, so the bar is the bf16 one
    np.testing.assert_allclose(a[:3][v[:3]], small.cpu().numpy()[v[:3]], rtol=0, atol=1e-2 * np.abs(b).max())


def test_pointnet_stash_is_bit_identical(cuda):
    """mpa_pointnet_forward with the large workspace (launches 4 and 5 resume from stashed operand
    tiles) and with the small one (every launch recomputes from the points) give the same bits."""
    from multi_part_assembly_b200 import _lib
    from multi_part_assembly_b200.kernels import _ptr_array
    from multi_part_assembly_b200.models import build_encoder
    enc = fill_params_(build_encoder('pointnet', 256), 3).to(cuda).train()
    convs = [m for m in enc.modules() if isinstance(m, torch.nn.Conv1d)]
    bns = [m for m in enc.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    g = torch.Generator().manual_seed(2)
    n, N = 37, 500
    x = (torch.rand(n, N, 3, generator=g) - 0.5).to(cuda).contiguous()
    valids = torch.ones(n, device=cuda)
    valids[5] = 0
    L = _lib.lib()
    w = [c.weight.detach().reshape(c.weight.shape[0], -1).float().contiguous() for c in convs]
    outs = []
    for big in (True, False):
        rm = [b.running_mean.clone() for b in bns]
        rv = [b.running_var.clone() for b in bns]
        nb = L.mpa_pointnet_workspace_bytes_n(n, N) if big else L.mpa_pointnet_workspace_bytes(n)
        ws = torch.empty(nb, dtype=torch.uint8, device=cuda)
        feats = torch.empty(n, 256, device=cuda)
        with torch.cuda.device(cuda):
            rc = L.mpa_pointnet_forward(_lib.ptr(x), _lib.ptr(valids), n, N, 256, _ptr_array(w),
                                        _ptr_array([b.weight.detach() for b in bns]),
                                        _ptr_array([b.bias.detach() for b in bns]), _ptr_array(rm), _ptr_array(rv),
                                        1, float(bns[0].eps), float(bns[0].momentum), _lib.ptr(feats), _lib.ptr(ws),
                                        nb, _lib.cuda_stream(cuda))
        _lib.check(rc, 'mpa_pointnet_forward')
        torch.cuda.synchronize()
        outs.append((feats.cpu(), [t.cpu() for t in rm + rv]))
    assert L.mpa_pointnet_workspace_bytes_n(n, N) > L.mpa_pointnet_workspace_bytes(n)
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)
