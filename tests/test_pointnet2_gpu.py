"""PointNet++ set-abstraction kernels and encoders vs the CPU oracle.

The reference's implementation is a CUDA-only extension (pointnet2_ops) that cannot run in
the build container, so there are no golden vectors from the reference itself: PARITY
UNPINNED.  The oracle (oracle/mpa_oracle.c `oracle_fps`, `oracle_ball_query`) restates the
reference kernels thread by thread, with the FMA contraction nvcc gives their distance
expression; the shared MLP is checked against stock torch layers on the same indices."""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from oracle.params import fill_params_

pytestmark = pytest.mark.gpu


def _clouds(B, n, seed, kind='uniform'):
    rng = np.random.default_rng(seed)
    x = (rng.random((B, n, 3)) - 0.5).astype(np.float32)
    if kind == 'duplicates':   # exact ties between candidates: the tie rule decides
        x[:, n // 2:] = x[:, :n - n // 2]
    elif kind == 'lattice':    # many equal distances
        x = (np.round(x * 8) / 8).astype(np.float32)
    elif kind == 'origin':     # points with |p|^2 <= 1e-3 are skipped by the reference
        x[:, ::3] *= 0.02
    return x


@pytest.mark.parametrize('B,n,m,kind', [(4, 1000, 512, 'uniform'), (3, 1000, 128, 'duplicates'),
                                        (2, 600, 512, 'lattice'), (3, 777, 200, 'origin'),
                                        (2, 512, 128, 'uniform'), (1, 100, 16, 'uniform')])
def test_fps_bit_exact(cuda, B, n, m, kind):
    from multi_part_assembly_b200 import kernels
    x = _clouds(B, n, n + m, kind)
    idx, new_xyz = kernels.furthest_point_sample(torch.from_numpy(x).to(cuda), m)
    want = oracle.furthest_point_sample(x, m)
    np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64), want)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), np.take_along_axis(x, want[..., None].repeat(3, -1), 1))


@pytest.mark.parametrize('B,n,m,radius,nsample', [(4, 1000, 512, 0.2, 64), (2, 1000, 512, 0.1, 16),
                                                  (3, 512, 128, 0.4, 128), (2, 300, 50, 0.01, 8)])
def test_ball_query_exact(cuda, B, n, m, radius, nsample):
    from multi_part_assembly_b200 import kernels
    x = _clouds(B, n, n + nsample)
    c = x[:, :m] + (0.3 if radius == 0.01 else 0.0)  # far centroids: balls without any point
    got = kernels.ball_query(radius, nsample, torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda))
    want = oracle.ball_query(x, c, radius, nsample)
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize('arch,feat,precision,bar', [('pointnet2_ssg', 256, 'fp32', 3e-4),
                                                     ('pointnet2_msg', 128, 'fp32', 3e-4),
                                                     ('pointnet2_ssg', 128, 'bf16', 6e-2)])
def test_pointnet2_encoder(cuda, arch, feat, precision, bar):
    """Whole encoder (train-mode BatchNorm): sampling / grouping indices vs the oracle, features
    and running statistics vs stock torch layers applied to the same groups, backward runs."""
    import copy
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models import build_encoder
    enc = fill_params_(build_encoder(arch, feat), 9).to(cuda).train()
    ref = copy.deepcopy(enc)
    x = _clouds(5, 1000, 3)
    xt = torch.from_numpy(x).to(cuda)
    kernels.set_precision(precision)
    kernels._POINTNET2_TRACE = trace = []
    try:
        out = enc(xt)
    finally:
        kernels.set_precision('auto')
        kernels._POINTNET2_TRACE = None
    assert out.shape == (5, feat) and torch.isfinite(out).all()
    # first level: sampling and ball queries against the oracle
    kinds = [k for k, _ in trace]
    assert kinds[0] == 'fps'
    fps0 = trace[0][1].cpu().numpy().astype(np.int64)
    np.testing.assert_array_equal(fps0, oracle.furthest_point_sample(x, 512))
    centroids = np.take_along_axis(x, fps0[..., None].repeat(3, -1), 1)
    sa0 = enc.SA_modules[0]
    for (kind, idx), radius, nsample in zip(trace[1:], sa0.radii, sa0.nsamples):
        assert kind == 'ball'
        np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64),
                                      oracle.ball_query(x, centroids, radius, nsample))
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want = kernels._pointnet2_torch(xt, ref, trace)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    err = float((out - want).abs().max() / want.abs().max())
    assert err < bar, err
    a, b = enc.SA_modules[1].mlps[0][1], ref.SA_modules[1].mlps[0][1]
    np.testing.assert_allclose(a.running_var.cpu().numpy(), b.running_var.cpu().numpy(),
                               rtol=20 * bar, atol=1e-5)
    assert int(a.num_batches_tracked) == 1
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.parameters())


def test_pointnet2_in_model(cuda):
    """configs/pn_transformer with encoder = 'pointnet2_ssg' steps end to end."""
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    from multi_part_assembly_b200.compat.lightning import Trainer
    cfg = get_cfg('pn_transformer')
    cfg.model.encoder = 'pointnet2_ssg'
    model = build_model(cfg).to(cuda).train()
    model.trainer = Trainer()
    batch = make_batch(2, P=20, N=1000, num_valid=[3, 2], seed=0, device=cuda)
    loss = model.training_step(batch, 0)
    loss.backward()
    assert torch.isfinite(loss).item()
