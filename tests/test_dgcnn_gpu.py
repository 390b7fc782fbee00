"""k-NN graph and EdgeConv kernels vs the CPU oracle / torch formulation."""
import os

import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from oracle.params import fill_params_

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('n,N,C,k', [(3, 64, 3, 20), (4, 1000, 3, 20), (2, 1000, 64, 20),
                                     (2, 777, 128, 20), (1, 1500, 64, 20), (2, 33, 5, 7),
                                     # tensor-core path corners: one / two k-blocks, N % 4 != 0 (scalar
                                     # epilogue), several row tiles with a ragged tail, k = 32, C > 128
                                     # (CUDA-core tile kernel), fewer than k column groups (overflow path)
                                     (3, 130, 72, 20), (2, 1023, 8, 32), (2, 257, 128, 5), (1, 300, 136, 20),
                                     (5, 200, 64, 20)])
def test_knn_bit_exact_vs_oracle(cuda, n, N, C, k):
    """north_star: bit-exact k-NN indices.  Contract = the sorted index set per
    row (EdgeConv only sees the set); here even the order matches the oracle's
    (score desc, index asc) selection."""
    from multi_part_assembly_b200 import kernels
    rng = np.random.default_rng(N + C)
    x = rng.standard_normal((n, N, C)).astype(np.float32)
    got = kernels.knn(torch.from_numpy(x).to(cuda), k).cpu().numpy().astype(np.int64)
    want = oracle.knn(np.ascontiguousarray(x.transpose(0, 2, 1)), k)
    np.testing.assert_array_equal(np.sort(got, -1), want)
    assert np.all(got[:, :, 0] == np.arange(N)[None])  # self is the best match


def test_knn_golden_reference_sets(cuda):
    """Sets produced by the reference's own knn (tests/golden/dgcnn.npz)."""
    from multi_part_assembly_b200 import kernels
    g = dict(np.load(os.path.join(GOLD, 'dgcnn.npz')))
    got = kernels.knn(torch.from_numpy(g['x']).to(cuda), 20).cpu().numpy()
    np.testing.assert_array_equal(np.sort(got, -1), g['knn_sorted'])


def test_knn_ties_duplicates(cuda):
    from multi_part_assembly_b200 import kernels
    rng = np.random.default_rng(0)
    base = rng.standard_normal((1, 40, 3)).astype(np.float32)
    x = np.concatenate([base, base, base], 1)  # every point three times
    got = kernels.knn(torch.from_numpy(x).to(cuda), 20).cpu().numpy().astype(np.int64)
    want = oracle.knn(np.ascontiguousarray(x.transpose(0, 2, 1)), 20)
    np.testing.assert_array_equal(np.sort(got, -1), want)


def test_edge_aggregate_vs_torch(cuda):
    from multi_part_assembly_b200 import kernels
    g = torch.Generator().manual_seed(0)
    n, N, Co, k = 3, 200, 64, 20
    uv = torch.randn(n * N, 2 * Co, generator=g).to(cuda)
    idx = torch.stack([torch.stack([torch.randperm(N, generator=g)[:k] for _ in range(N)])
                       for _ in range(n)]).int().to(cuda)
    ymax, ymin, sums = kernels.edge_aggregate(uv, idx, n, N, Co, k)
    u = uv[:, :Co].view(n, N, Co)
    v = uv[:, Co:].view(n, N, 1, Co)
    y = torch.gather(u.unsqueeze(1).expand(n, N, N, Co), 2,
                     idx.long().unsqueeze(-1).expand(n, N, k, Co)) + v
    np.testing.assert_array_equal(ymax.view(n, N, Co).cpu().numpy(), y.max(2)[0].cpu().numpy())
    np.testing.assert_array_equal(ymin.view(n, N, Co).cpu().numpy(), y.min(2)[0].cpu().numpy())
    # fp32 per-CTA partial sums: tolerance relative to the sum of magnitudes
    scale = float(y.abs().double().sum((0, 1, 2)).max())
    np.testing.assert_allclose(sums[:, 0].cpu().numpy(), y.double().sum((0, 1, 2)).cpu().numpy(),
                               rtol=0, atol=1e-6 * scale)
    np.testing.assert_allclose(sums[:, 1].cpu().numpy(), (y.double()**2).sum((0, 1, 2)).cpu().numpy(), rtol=1e-5)


def test_dgcnn_bf16_mode(cuda):
    """bf16 tensor-core GEMMs inside DGCNN vs the reference's fp32 golden output: same
    first-layer graph, features within bf16 operand rounding (5e-2 of the largest feature;
    tests/test_bench_config_gpu.py holds the N = 1000 version with the graph check)."""
    from multi_part_assembly_b200 import kernels
    from multi_part_assembly_b200.models import build_encoder
    g = dict(np.load(os.path.join(GOLD, 'dgcnn.npz')))
    enc = fill_params_(build_encoder('dgcnn', 128), 5).to(cuda).train()
    kernels.set_precision('bf16')
    try:
        with torch.no_grad():
            out = enc(torch.from_numpy(g['x']).to(cuda)).detach().float().cpu().numpy()
    finally:
        kernels.set_precision('auto')
    assert np.isfinite(out).all()
    err = np.abs(out - g['out_train']).max() / np.abs(g['out_train']).max()
    assert err < 5e-2, err


def test_knn_clustered_in_few_lanes(cuda):
    """Neighbours that all sit at indices j % 32 in {0,1,2}: the lane-maximum threshold of
    the tiled kernel then lets more than 64 candidates through and the slow exact path
    (warp arg-max rounds) must take over.  Still bit-exact."""
    from multi_part_assembly_b200 import kernels
    j = np.arange(1024)
    near = (j % 32) < 3
    x = np.where(near, j * 1e-3, 1000.0 + j).astype(np.float32).reshape(1, 1024, 1)
    x = np.concatenate([x, x[:, ::-1]], 0)  # second part: same values, reversed order
    got = kernels.knn(torch.from_numpy(np.ascontiguousarray(x)).to(cuda), 20).cpu().numpy().astype(np.int64)
    want = oracle.knn(np.ascontiguousarray(x.transpose(0, 2, 1)), 20)
    np.testing.assert_array_equal(np.sort(got, -1), want)


@pytest.mark.parametrize('global_feat', [True, False])
def test_dgcnn_valids_mask_equals_compaction(cuda, global_feat):
    """Masking padded parts on the device (`valids`) gives what the reference's gather ->
    encoder -> scatter gives (models/dgl/network.py:90-99): same features on the valid parts,
    zeros on the padding, same BatchNorm running statistics -- with no host sync, so the step
    can be captured in a CUDA graph."""
    import copy
    from multi_part_assembly_b200.models import build_encoder
    enc = fill_params_(build_encoder('dgcnn', 128, global_feat=global_feat), 5).to(cuda).train()
    ref = copy.deepcopy(enc)
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(7, 300, 3, generator=g) - 0.5).to(cuda)
    valids = torch.tensor([1, 1, 0, 1, 0, 0, 1.], device=cuda)
    x = x * valids.view(-1, 1, 1)  # padded parts are all-zero clouds, as the dataset pads them
    keep = valids.bool()
    with torch.no_grad():
        got = enc(x, valids=valids)
        want = ref(x[keep])
    np.testing.assert_allclose(got[keep].cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5)
    assert torch.all(got[~keep] == 0)
    for name in ('bn1', 'bn4', 'bn5'):
        a, b = getattr(enc, name), getattr(ref, name)
        np.testing.assert_allclose(a.running_mean.cpu().numpy(), b.running_mean.cpu().numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(a.running_var.cpu().numpy(), b.running_var.cpu().numpy(), rtol=1e-5, atol=1e-7)


def test_knn_skips_padded_parts(cuda):
    """`valids`: padded parts keep zero rows, the others are unaffected (tensor-core path: the
    Gram kernel skips the tiles of padded items)."""
    from multi_part_assembly_b200 import kernels
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6, 300, 64)).astype(np.float32)
    valids = torch.tensor([1, 0, 1, 1, 0, 1.], device=cuda)
    got = kernels.knn(torch.from_numpy(x).to(cuda), 20, valids=valids).cpu().numpy().astype(np.int64)
    want = oracle.knn(np.ascontiguousarray(x.transpose(0, 2, 1)), 20)
    keep = valids.cpu().numpy() != 0
    np.testing.assert_array_equal(np.sort(got[keep], -1), want[keep])
    assert (got[~keep] == 0).all()
