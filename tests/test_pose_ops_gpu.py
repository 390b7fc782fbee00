"""SE(3) kernel and the fused pose-Chamfer losses vs the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle

pytestmark = pytest.mark.gpu


def _rand_quat(rng, shape, unit=True):
    q = rng.standard_normal(shape + (4, )).astype(np.float32)
    if unit:
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
    return q.astype(np.float32)


@pytest.mark.parametrize('N', [1, 7, 100, 1000])
@pytest.mark.parametrize('with_trans', [False, True])
def test_se3_bit_exact(cuda, N, with_trans):
    from multi_part_assembly_b200.utils import qrot, qtransform
    rng = np.random.default_rng(N)
    q = _rand_quat(rng, (3, 5), unit=(N != 7))  # non-unit q scales by |q|^2 like the reference
    t = rng.standard_normal((3, 5, 3)).astype(np.float32)
    v = rng.standard_normal((3, 5, N, 3)).astype(np.float32)
    tq, tt, tv = [torch.from_numpy(a).to(cuda) for a in (q, t, v)]
    got = (qtransform(tt, tq, tv) if with_trans else qrot(tq, tv)).cpu().numpy()
    want = oracle.se3_transform(q, t if with_trans else None, v)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


def test_se3_matches_scipy(cuda):
    """Independent pin of the pytorch3d convention (real part first, active
    rotation) -- pytorch3d itself is absent, so scipy is the third opinion."""
    from scipy.spatial.transform import Rotation as R
    from multi_part_assembly_b200.utils import qrot
    rng = np.random.default_rng(0)
    q = _rand_quat(rng, (6, ))
    v = rng.standard_normal((6, 50, 3)).astype(np.float32)
    got = qrot(torch.from_numpy(q).to(cuda), torch.from_numpy(v).to(cuda)).cpu().numpy()
    want = np.stack([R.from_quat(q[i, [1, 2, 3, 0]]).apply(v[i]) for i in range(6)])
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_se3_backward(cuda):
    from multi_part_assembly_b200.utils import qtransform
    from oracle import torch_ref
    rng = np.random.default_rng(1)
    q = torch.from_numpy(_rand_quat(rng, (4, ), unit=False))
    t = torch.from_numpy(rng.standard_normal((4, 3)).astype(np.float32))
    v = torch.from_numpy(rng.standard_normal((4, 33, 3)).astype(np.float32))
    w = torch.from_numpy(rng.standard_normal((4, 33, 3)).astype(np.float32))
    ref = [x.clone().double().requires_grad_() for x in (t, q, v)]
    (torch_ref.qtransform(*ref) * w.double()).sum().backward()
    got = [x.clone().to(cuda).requires_grad_() for x in (t, q, v)]
    (qtransform(*got) * w.to(cuda)).sum().backward()
    for g, r in zip(got, ref):
        np.testing.assert_allclose(g.grad.cpu().numpy(), r.grad.numpy(), rtol=2e-5, atol=2e-5)


def _pose_inputs(B, P, N, n_valid, seed):
    rng = np.random.default_rng(seed)
    pts = (rng.random((B, P, N, 3)) - 0.5).astype(np.float32)
    valids = np.zeros((B, P), np.float32)
    for b in range(B):
        valids[b, :n_valid[b % len(n_valid)]] = 1
    pts[valids == 0] = 0
    q1 = _rand_quat(rng, (B, P)); q2 = _rand_quat(rng, (B, P))
    q2[valids == 0] = [1, 0, 0, 0]  # Rotation3D rewrites padded GT quats to identity
    t1 = (rng.standard_normal((B, P, 3)) * 0.3).astype(np.float32)
    t2 = ((rng.random((B, P, 3)) - 0.5)).astype(np.float32)
    t2[valids == 0] = 0
    return pts, valids, q1, t1, q2, t2


def _pose_chamfer_gpu(pts, valids, q1, t1, q2, t2, mode, dev):
    from multi_part_assembly_b200 import _lib
    B, P, N, _ = pts.shape
    tt = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    d = dict(pts=tt(pts), q1=tt(q1), t1=tt(t1), q2=tt(q2), t2=tt(t2), v=tt(valids))
    out = dict(d1=torch.full((B, P, N), -7., device=dev), d2=torch.full((B, P, N), -7., device=dev),
               i1=torch.full((B, P, N), -7, dtype=torch.int32, device=dev),
               i2=torch.full((B, P, N), -7, dtype=torch.int32, device=dev),
               p1=torch.empty(B, P, N, 3, device=dev), p2=torch.empty(B, P, N, 3, device=dev))
    L = _lib.lib()
    rc = L.mpa_pose_chamfer(
        _lib.ptr(d['pts']), _lib.ptr(d['q1']), _lib.ptr(d['t1']), _lib.ptr(d['q2']),
        _lib.ptr(d['t2']), _lib.ptr(d['v']), B, P, N, mode, _lib.ptr(out['d1']),
        _lib.ptr(out['i1']), _lib.ptr(out['d2']), _lib.ptr(out['i2']), _lib.ptr(out['p1']),
        _lib.ptr(out['p2']), None, 0, _lib.cuda_stream(dev))
    _lib.check(rc, 'mpa_pose_chamfer')
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize('B,P,N,n_valid', [(2, 4, 100, [4, 2]), (3, 20, 1000, [20, 7, 2]),
                                           (1, 3, 37, [1])])
def test_pose_chamfer_part_mode(cuda, B, P, N, n_valid):
    """rot_points_cd_loss core (loss.py:125-131): rot_pc x2 + per-part Chamfer."""
    pts, valids, q1, t1, q2, t2 = _pose_inputs(B, P, N, n_valid, 11)
    got = _pose_chamfer_gpu(pts, valids, q1, None, q2, None, 0, cuda)
    p1 = oracle.se3_transform(q1, None, pts); p2 = oracle.se3_transform(q2, None, pts)
    np.testing.assert_array_equal(got['p1'], p1); np.testing.assert_array_equal(got['p2'], p2)
    e1, j1, e2, j2 = oracle.chamfer_forward(p1.reshape(B * P, N, 3), p2.reshape(B * P, N, 3))
    m = valids.reshape(-1) == 1
    for g, e in ((got['d1'], e1), (got['d2'], e2), (got['i1'], j1), (got['i2'], j2)):
        g = g.reshape(B * P, N)
        np.testing.assert_array_equal(g[m], e[m].astype(g.dtype))
    assert np.all(got['d1'].reshape(B * P, N)[~m] == 0) and np.all(got['i1'].reshape(B * P, N)[~m] == -1)


@pytest.mark.parametrize('B,P,N,n_valid', [(2, 4, 100, [4, 2]), (3, 20, 1000, [20, 7, 2]),
                                           (2, 5, 64, [5, 1])])
def test_pose_chamfer_shape_mode(cuda, B, P, N, n_valid):
    """shape_cd_loss core (loss.py:170-182): 1e3 fill, transform_pc x2, Chamfer
    over the P*N points of a shape; padded target parts stay candidates."""
    pts, valids, q1, t1, q2, t2 = _pose_inputs(B, P, N, n_valid, 13)
    got = _pose_chamfer_gpu(pts, valids, q1, t1, q2, t2, 1, cuda)
    filled = pts.copy(); filled[valids == 0] = 1e3
    p1 = oracle.se3_transform(q1, t1, filled); p2 = oracle.se3_transform(q2, t2, filled)
    np.testing.assert_array_equal(got['p1'], p1); np.testing.assert_array_equal(got['p2'], p2)
    e1, j1, e2, j2 = oracle.chamfer_forward(p1.reshape(B, P * N, 3), p2.reshape(B, P * N, 3))
    m = np.repeat(valids.reshape(B, P, 1), N, 2).reshape(B, P * N) == 1
    for g, e in ((got['d1'], e1), (got['d2'], e2), (got['i1'], j1), (got['i2'], j2)):
        g = g.reshape(B, P * N)
        np.testing.assert_array_equal(g[m], e[m].astype(g.dtype))
        assert np.all(g[~m] == (0 if g.dtype == np.float32 else -1))


@pytest.mark.parametrize('spread', [0.0, 0.02, 3.0])
def test_pose_chamfer_blob_vs_spread(cuda, spread):
    """The distribution the benchmark feeds the search: an untrained model puts
    every part under nearly the same pose (a dense blob of 20 000 points) while the ground truth
    is spread over a box several times larger, so ~40 % of the ground-truth queries lie far
    outside the target grid and take the pyramid descent.  spread = 3: clouds that do not
    overlap at all (every query is a far query).  Bit-exact vs the brute-force oracle."""
    B, P, N = 2, 20, 1000
    pts, valids, q1, t1, q2, t2 = _pose_inputs(B, P, N, [20, 13], 23)
    rng = np.random.default_rng(5)
    base = _rand_quat(rng, (B, 1))
    q1 = base + 0.03 * rng.standard_normal((B, P, 4)).astype(np.float32)
    q1 = (q1 / np.linalg.norm(q1, axis=-1, keepdims=True)).astype(np.float32)
    t1 = (0.2 + 0.02 * rng.standard_normal((B, P, 3)) + spread).astype(np.float32)
    if spread == 0.0:
        t1[:] = 0.25  # all predicted translations identical
    got = _pose_chamfer_gpu(pts, valids, q1, t1, q2, t2, 1, cuda)
    filled = pts.copy(); filled[valids == 0] = 1e3
    p1 = oracle.se3_transform(q1, t1, filled); p2 = oracle.se3_transform(q2, t2, filled)
    e1, j1, e2, j2 = oracle.chamfer_forward(p1.reshape(B, P * N, 3), p2.reshape(B, P * N, 3))
    m = np.repeat(valids.reshape(B, P, 1), N, 2).reshape(B, P * N) == 1
    for g, e in ((got['d1'], e1), (got['d2'], e2), (got['i1'], j1), (got['i2'], j2)):
        g = g.reshape(B, P * N)
        np.testing.assert_array_equal(g[m], e[m].astype(g.dtype))


def test_shape_mode_far_part_can_win(cuda):
    """A valid query that sits next to a padded part's 1e3 point must report it,
    exactly as the reference's brute force would."""
    B, P, N = 1, 3, 64
    pts, valids, q1, t1, q2, t2 = _pose_inputs(B, P, N, [2], 17)
    q1[:] = [1, 0, 0, 0]; q2[:] = [1, 0, 0, 0]
    t1[0, 0] = 1e3  # part 0 of cloud 1 moved next to the padded fill point of cloud 2
    got = _pose_chamfer_gpu(pts, valids, q1, t1, q2, t2, 1, cuda)
    filled = pts.copy(); filled[valids == 0] = 1e3
    p1 = oracle.se3_transform(q1, t1, filled); p2 = oracle.se3_transform(q2, t2, filled)
    e1, j1, _, _ = oracle.chamfer_forward(p1.reshape(B, P * N, 3), p2.reshape(B, P * N, 3))
    assert np.all(j1[0, :N] == 2 * N)  # lowest index of the padded part
    np.testing.assert_array_equal(got['i1'].reshape(B, -1)[0, :N], j1[0, :N])
    np.testing.assert_array_equal(got['d1'].reshape(B, -1)[0, :N], e1[0, :N])


def test_lsap_kernel_matches_scipy(cuda):
    """Batched assignment kernel vs SciPy's linear_sum_assignment (the reference's call,
    base_model.py:175-176): identical column assignment, including tie-heavy matrices."""
    from scipy.optimize import linear_sum_assignment
    from multi_part_assembly_b200 import kernels
    rng = np.random.default_rng(1)
    mats = []
    for trial in range(400):
        n = int(rng.integers(1, 21))
        kind = trial % 4
        if kind == 0:
            c = rng.random((n, n))
        elif kind == 1:
            c = rng.integers(0, 4, (n, n))
        elif kind == 2:
            c = np.full((n, n), 0.5)
        else:
            c = rng.random((n, n)) * 1e-3 + rng.integers(0, 2, (n, n))
        mats.append(c.astype(np.float32))
    mats.append(rng.random((64, 64)).astype(np.float32))
    got = kernels.lsap_batched([torch.from_numpy(m).to(cuda) for m in mats])
    for m, g in zip(mats, got):
        np.testing.assert_array_equal(g.cpu().numpy(), linear_sum_assignment(m)[1])


@pytest.mark.gpu
def test_rotation3d_zero_quat_native(cuda):
    """Rotation3D's constructor on CUDA (one native launch) vs the reference rule
    (rotation.py:121-128): zero quaternions -> identity, the rest bit-identical."""
    from multi_part_assembly_b200.utils.rotation import Rotation3D
    g = torch.Generator().manual_seed(3)
    q = torch.nn.functional.normalize(torch.randn(7, 20, 4, generator=g), dim=-1)
    q[1, 5:] = 0.
    q[4] = 0.
    q[6, 0] = torch.tensor([0.3, 0.2, 0.1, 0.1])  # norm < 0.5 -> identity too
    want = Rotation3D(q.clone(), 'quat').rot                # CPU: the torch formulation
    got = Rotation3D(q.to(cuda), 'quat').rot
    assert torch.equal(got.cpu(), want)
    # a tensor that needs a gradient keeps the differentiable torch path
    qg = q.to(cuda).requires_grad_(True)
    r = Rotation3D(qg, 'quat').rot
    assert r.requires_grad and torch.equal(r.detach().cpu(), want)
