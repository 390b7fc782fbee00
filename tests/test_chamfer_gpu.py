"""Parity of the sm_100a Chamfer / SE(3) kernels (through the C ABI) against the
CPU oracle.  Bar: distances bit-exact, indices exact (north_star: Chamfer within
1e-5 rel -- we hold the stronger bit-exact bar because the kernels reproduce the
reference kernel's rounding sequence)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import cpu as oracle

pytestmark = pytest.mark.gpu

ALGOS = {'auto': 0, 'brute': 1, 'grid': 2}


def _fwd(x1, x2, algo, dev):
    from multi_part_assembly_b200.utils.chamfer import chamfer_forward
    out = chamfer_forward(torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev),
                          algo=ALGOS[algo])
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def _check_exact(x1, x2, algo, dev):
    d1, i1, d2, i2 = _fwd(x1, x2, algo, dev)
    e1, j1, e2, j2 = oracle.chamfer_forward(x1, x2)
    np.testing.assert_array_equal(d1.view(np.uint32), e1.view(np.uint32))
    np.testing.assert_array_equal(d2.view(np.uint32), e2.view(np.uint32))
    np.testing.assert_array_equal(i1, j1)
    np.testing.assert_array_equal(i2, j2)


@pytest.mark.parametrize('algo', ['auto', 'brute', 'grid'])
def test_config_a(cuda, algo):
    """BASELINE config A: 2 parts x 500 pts, torch.rand, seed 0."""
    torch.manual_seed(0)
    x1 = torch.rand(2, 500, 3).numpy()
    x2 = torch.rand(2, 500, 3).numpy()
    _check_exact(x1, x2, algo, cuda)
    # and against the reference test's own brute-force definition (atol 1e-6,
    # test_chamfer.py:72-76)
    d1, i1, d2, i2 = _fwd(x1, x2, algo, cuda)
    D = ((x1[:, :, None].astype(np.float64) - x2[:, None].astype(np.float64))**2).sum(-1)
    np.testing.assert_allclose(d1, D.min(2), atol=1e-6)
    np.testing.assert_array_equal(i1, D.argmin(2))
    np.testing.assert_array_equal(i2, D.argmin(1))


@pytest.mark.parametrize('algo', ['brute', 'grid'])
@pytest.mark.parametrize('shape', [(1, 1, 1), (3, 17, 1000), (2, 1000, 37), (4, 513, 769),
                                   (1, 2049, 4097), (5, 64, 64)])
def test_ragged_sizes(cuda, algo, shape):
    B, n1, n2 = shape
    rng = np.random.default_rng(n1 * 7 + n2)
    x1 = rng.standard_normal((B, n1, 3)).astype(np.float32)
    x2 = (rng.standard_normal((B, n2, 3)) * 0.5 + 0.3).astype(np.float32)
    _check_exact(x1, x2, algo, cuda)


@pytest.mark.parametrize('algo', ['brute', 'grid'])
def test_ties_take_lowest_index(cuda, algo):
    """Duplicated points: chamfer_kernel.cu:82 keeps the first minimum."""
    rng = np.random.default_rng(3)
    base = rng.random((2, 200, 3)).astype(np.float32)
    x2 = np.concatenate([base, base[:, ::-1], base], axis=1)  # every point 3x
    x1 = np.concatenate([base[:, :50], rng.random((2, 300, 3)).astype(np.float32)], axis=1)
    _check_exact(x1, x2, algo, cuda)
    # lattice data: many exactly equal distances
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing='ij'), -1).reshape(1, -1, 3)
    x1 = (g * 0.25).astype(np.float32)
    x2 = (g[:, ::-1] * 0.25 + 0.125).astype(np.float32)
    _check_exact(x1, x2, algo, cuda)


@pytest.mark.parametrize('algo', ['brute', 'grid'])
def test_degenerate_clouds(cuda, algo):
    rng = np.random.default_rng(5)
    # all points identical / planar / collinear / far outliers (the 1e3 fill of
    # shape_cd_loss, loss.py:175) / clustered
    same = np.full((2, 700, 3), 0.25, np.float32)
    rnd = rng.random((2, 600, 3)).astype(np.float32)
    _check_exact(same, rnd, algo, cuda)
    _check_exact(same, same.copy(), algo, cuda)
    planar = rnd.copy(); planar[..., 2] = 0.5
    _check_exact(planar, rnd, algo, cuda)
    line = rnd.copy(); line[..., 1:] = 0.0
    _check_exact(line, planar, algo, cuda)
    out = rnd.copy(); out[:, 300:] = 1e3
    out2 = rng.random((2, 650, 3)).astype(np.float32); out2[:, 100:200] = 1e3
    out2[:, 200:250] += 1e3
    _check_exact(out, out2, algo, cuda)
    clus = (rng.integers(0, 4, (2, 900, 1)) * 10.0 + rng.standard_normal((2, 900, 3)) * 0.01)
    _check_exact(clus.astype(np.float32), rnd * 30, algo, cuda)


def test_grid_equals_brute_full_size(cuda):
    """BASELINE cfg C shape-level size [32, 20000, 3]: the two algorithms agree
    bit for bit, and a subsample agrees with the oracle."""
    g = torch.Generator().manual_seed(1)
    x1 = (torch.rand(8, 20000, 3, generator=g) - 0.5) * 2
    x2 = (torch.rand(8, 20000, 3, generator=g) - 0.5) * 2
    a = _fwd(x1.numpy(), x2.numpy(), 'grid', cuda)
    b = _fwd(x1.numpy(), x2.numpy(), 'brute', cuda)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)
    e1, j1, e2, j2 = oracle.chamfer_forward(x1[:1].numpy(), x2[:1].numpy())
    np.testing.assert_array_equal(a[0][:1], e1)
    np.testing.assert_array_equal(a[1][:1], j1)
    np.testing.assert_array_equal(a[3][:1], j2)


@pytest.mark.parametrize('n,offset', [(20000, 0.0), (20000, 1.5), (1000, 0.0), (1000, 4.0), (3000, 0.7)])
def test_grid_far_queries(cuda, n, offset):
    """Dense blob vs. a cloud spread over a much larger box (an untrained model's
    assembly vs. the ground truth): most queries of one direction lie far outside the
    target cloud and take the two-level block search; results must still equal the
    brute force bit for bit."""
    rng = np.random.default_rng(n + int(offset * 10))
    blob = (rng.random((3, n, 3)) - 0.5).astype(np.float32)
    parts = (rng.random((3, 20, 1, 3)) - 0.5) * 2.0
    spread = ((rng.random((3, 20, n // 20, 3)) - 0.5) + parts).reshape(3, -1, 3).astype(np.float32)
    spread += np.float32(offset)
    spread[0, :5] = 50.0  # a few extreme outliers stretch the bounding box
    a = _fwd(blob, spread, 'grid', cuda)
    b = _fwd(blob, spread, 'brute', cuda)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)


def test_empty_inputs(cuda):
    from multi_part_assembly_b200.utils.chamfer import chamfer_forward
    d1, i1, d2, i2 = chamfer_forward(torch.zeros(0, 10, 3, device=cuda),
                                     torch.zeros(0, 7, 3, device=cuda))
    assert d1.shape == (0, 10) and i2.shape == (0, 7)
    # no targets: the reference writes its init values 1e32 / -1 (chamfer_kernel.cu:60-61)
    d1, i1, d2, i2 = chamfer_forward(torch.rand(2, 5, 3, device=cuda),
                                     torch.zeros(2, 0, 3, device=cuda))
    assert d2.shape == (2, 0)
    assert torch.all(d1 == 1e32) and torch.all(i1 == -1)


def test_rejects_cpu_and_bad_shapes(cuda):
    from multi_part_assembly_b200.utils.chamfer import chamfer_forward
    with pytest.raises(RuntimeError):
        chamfer_forward(torch.rand(1, 4, 3), torch.rand(1, 4, 3))  # CPU tensors
    with pytest.raises(RuntimeError):
        chamfer_forward(torch.rand(1, 4, 2, device=cuda), torch.rand(1, 4, 3, device=cuda))
    with pytest.raises(RuntimeError):
        chamfer_forward(torch.rand(2, 4, 3, device=cuda), torch.rand(1, 4, 3, device=cuda))


def test_backward_matches_oracle(cuda):
    from multi_part_assembly_b200.utils.chamfer import chamfer_distance
    torch.manual_seed(0)
    x1 = torch.rand(2, 64, 3)
    x2 = torch.rand(2, 80, 3)
    g1 = torch.rand(2, 64)
    g2 = torch.rand(2, 80)
    a = x1.to(cuda).requires_grad_()
    b = x2.to(cuda).requires_grad_()
    d1, d2 = chamfer_distance(a, b)
    ((d1 * g1.to(cuda)).sum() + (d2 * g2.to(cuda)).sum()).backward()
    _, i1, _, i2 = oracle.chamfer_forward(x1.numpy(), x2.numpy())
    e1, e2 = oracle.chamfer_backward(g1.numpy(), g2.numpy(), x1.numpy(), x2.numpy(), i1, i2)
    np.testing.assert_allclose(a.grad.cpu().numpy(), e1, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(b.grad.cpu().numpy(), e2, rtol=1e-5, atol=1e-6)


def test_backward_finite_difference(cuda):
    """Stand-in for the reference's float64 gradcheck (test_chamfer.py:92-101):
    central differences in float64 on the CPU oracle vs the CUDA gradient."""
    from multi_part_assembly_b200.utils.chamfer import chamfer_distance
    rng = np.random.default_rng(0)
    x1 = rng.random((1, 16, 3)).astype(np.float32)
    x2 = rng.random((1, 16, 3)).astype(np.float32)

    def f(a, b):
        D = ((a[:, :, None].astype(np.float64) - b[:, None].astype(np.float64))**2).sum(-1)
        return D.min(2).sum() + D.min(1).sum()

    a = torch.from_numpy(x1).to(cuda).requires_grad_()
    b = torch.from_numpy(x2).to(cuda).requires_grad_()
    d1, d2 = chamfer_distance(a, b)
    (d1.sum() + d2.sum()).backward()
    eps = 1e-4
    num = np.zeros_like(x1, dtype=np.float64)
    for i in range(16):
        for c in range(3):
            p = x1.astype(np.float64).copy(); p[0, i, c] += eps
            m = x1.astype(np.float64).copy(); m[0, i, c] -= eps
            num[0, i, c] = (f(p, x2) - f(m, x2)) / (2 * eps)
    np.testing.assert_allclose(a.grad.cpu().numpy(), num, rtol=1e-3, atol=1e-3)


def test_host_entry_point(cuda):
    """C ABI with HOST buffers (the e2e leg of bench.py)."""
    from multi_part_assembly_b200 import _lib
    rng = np.random.default_rng(9)
    x1 = rng.random((3, 1000, 3)).astype(np.float32)
    x2 = rng.random((3, 800, 3)).astype(np.float32)
    d1 = np.empty((3, 1000), np.float32); i1 = np.empty((3, 1000), np.int64)
    d2 = np.empty((3, 800), np.float32); i2 = np.empty((3, 800), np.int64)
    torch.cuda.init(); torch.zeros(1, device=cuda)
    rc = _lib.lib().mpa_chamfer_forward_host(
        x1.ctypes.data, x2.ctypes.data, 3, 1000, 800, d1.ctypes.data, i1.ctypes.data,
        d2.ctypes.data, i2.ctypes.data, 0, None)
    _lib.check(rc, 'mpa_chamfer_forward_host')
    e1, j1, e2, j2 = oracle.chamfer_forward(x1, x2)
    np.testing.assert_array_equal(d1, e1); np.testing.assert_array_equal(i1, j1)
    np.testing.assert_array_equal(d2, e2); np.testing.assert_array_equal(i2, j2)


def test_error_reporting(cuda):
    from multi_part_assembly_b200 import _lib
    L = _lib.lib()
    rc = L.mpa_chamfer_forward(None, None, 1, 4, 4, None, None, None, None, 0, None, 0, None)
    assert rc == -1 and b'null pointer' in L.mpa_last_error()
    rc = L.mpa_chamfer_forward(None, None, 1, 4, 4, None, None, None, None, 7, None, 0, None)
    assert rc == -1 and b'bad algo' in L.mpa_last_error()
