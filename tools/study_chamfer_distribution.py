"""CPU study (numpy) of the Chamfer search on the bench-like distribution: an untrained
model's assembly (all parts under nearly the same pose -> a dense blob) vs the ground truth
(parts spread by U(-0.5,0.5)^3 translations, random rotations).  Counts, per direction,
how many targets lie in the 3x3x3 block of a query's cell on the target grid, how many of
those survive the per-row pruning by the running best, and how many queries cannot be
proven complete there (they take the block search).  Informs DESIGN.md 4a."""
import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation as R

rng = np.random.default_rng(0)
P, N, OCC = 20, 1000, 3.0
parts = rng.random((P, N, 3)) - 0.5
parts -= parts.mean(1, keepdims=True)
gt = np.concatenate([R.random(random_state=i).apply(parts[i]) + (rng.random(3) - 0.5) for i in range(P)])
q0 = R.random(random_state=99)
pred = np.concatenate([(q0 * R.from_rotvec(0.05 * rng.standard_normal(3))).apply(parts[i]) +
                       0.2 + 0.02 * rng.standard_normal(3) for i in range(P)])


def study(Q, T, name):
    lo, hi = T.min(0), T.max(0)
    h = ((hi - lo).prod() * OCC / len(T)) ** (1 / 3)
    dims = np.floor((hi - lo) / h).astype(int) + 1
    cell = np.clip(np.floor((T - lo) / h).astype(int), 0, dims - 1)
    key = (cell[:, 2] * dims[1] + cell[:, 1]) * dims[0] + cell[:, 0]
    counts = np.bincount(key, minlength=dims.prod()).reshape(dims[2], dims[1], dims[0])
    qc = np.clip(np.floor((Q - lo) / h).astype(int), 0, dims - 1)
    pad = np.pad(counts, 1)
    blk = np.zeros(len(Q), int)
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                blk += pad[qc[:, 2] + dz, qc[:, 1] + dy, qc[:, 0] + dx]
    own = counts[qc[:, 2], qc[:, 1], qc[:, 0]]
    d, _ = cKDTree(T).query(Q)
    # distance to the faces of the 3x3x3 block that have cells beyond them
    lo_face = lo + (qc - 1) * h
    hi_face = lo + (qc + 2) * h
    bound = np.full(len(Q), np.inf)
    for ax in range(3):
        has_lo = qc[:, ax] - 1 > 0
        has_hi = qc[:, ax] + 1 < dims[ax] - 1
        bound = np.where(has_lo, np.minimum(bound, Q[:, ax] - lo_face[:, ax]), bound)
        bound = np.where(has_hi, np.minimum(bound, hi_face[:, ax] - Q[:, ax]), bound)
    hard = d >= bound
    outside = ((Q < lo) | (Q > hi)).any(1)
    q = lambda a: np.percentile(a, [50, 90, 99, 100]).round(1)
    print(f'{name}: grid {dims} h={h:.3f}; targets in own cell p50/p90/p99/max {q(own)}, in 3x3x3 block {q(blk)} '
          f'(mean {blk.mean():.0f}); NN distance / h p50/p90/p99 {np.percentile(d / h, [50, 90, 99]).round(2)}; '
          f'queries outside the target bbox {outside.mean():.0%}; not provable in the block {hard.mean():.0%}')
    # imbalance inside a warp of 32 consecutive queries in query-cell order
    lo_q, hi_q = Q.min(0), Q.max(0)
    hq = ((hi_q - lo_q).prod() * OCC / len(Q)) ** (1 / 3)
    dq = np.floor((hi_q - lo_q) / hq).astype(int) + 1
    cq = np.clip(np.floor((Q - lo_q) / hq).astype(int), 0, dq - 1)
    order = np.argsort((cq[:, 2] * dq[1] + cq[:, 1]) * dq[0] + cq[:, 0], kind='stable')
    b = blk[order][: len(Q) // 32 * 32].reshape(-1, 32)
    hw = hard[order][: len(Q) // 32 * 32].reshape(-1, 32)
    print(f'    per warp: max/mean of the block population {np.mean(b.max(1) / np.maximum(b.mean(1), 1)):.2f}; '
          f'warps with at least one hard query {np.mean(hw.any(1)):.0%}, hard lanes in those {hw.sum() / max(hw.any(1).sum(), 1):.1f}/32')


study(pred, gt, 'direction A (assembly -> ground truth)')
study(gt, pred, 'direction B (ground truth -> assembly)')
