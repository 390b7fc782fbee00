"""Evaluation metrics, reference API of utils/eval_utils.py."""
import itertools

import torch

from .loss import _valid_mean, pose_chamfer, CD_PART
from .chamfer import chamfer_distance
from .rotation import Rotation3D
from .transforms import transform_pc


@torch.no_grad()
def calc_part_acc(pts, trans1, trans2, rot1, rot2, valids):
    """Part accuracy: share of valid parts whose Chamfer distance between the
    two posed copies is below 0.01 (reference :13-46).  Returns [B]."""
    B, P = pts.shape[:2]
    if rot1.rot_type == 'quat' and rot2.rot_type == 'quat':
        dist1, dist2, _, _ = pose_chamfer(pts, trans1, trans2, rot1.rot,
                                          rot2.rot, valids, CD_PART)
        cd = dist1.mean(-1) + dist2.mean(-1)
    else:
        pts1 = transform_pc(trans1, rot1, pts).flatten(0, 1)
        pts2 = transform_pc(trans2, rot2, pts).flatten(0, 1)
        dist1, dist2 = chamfer_distance(pts1, pts2)
        cd = (dist1.mean(1) + dist2.mean(1)).view(B, P)
    cd = cd.type_as(pts)
    acc = (cd < 0.01) & (valids == 1)
    return acc.sum(-1) / (valids == 1).sum(-1)


def get_sym_point(point, x, y, z):
    """Mirror `point` along the flagged axes."""
    sign = point.new_tensor([-1. if x == 1 else 1., -1. if y == 1 else 1.,
                             -1. if z == 1 else 1.])
    return point * sign


def get_sym_point_list(point, sym=None):
    """All mirror images of `point` for the symmetry flags `sym` (default all
    three axes), in the reference's x-major order (reference :126-142)."""
    if sym is None:
        sym = [1, 1, 1]
    elif not isinstance(sym, (list, tuple)):
        sym = sym.tolist()
    sym = [int(i) for i in sym]
    return [get_sym_point(point, x, y, z) for x, y, z in itertools.product(
        range(sym[0] + 1), range(sym[1] + 1), range(sym[2] + 1))]


@torch.no_grad()
def calc_connectivity_acc(trans, rot, contact_points):
    """Connectivity accuracy (reference :50-110): for every contacting part
    pair, the min distance between the two posed contact points (over their
    mirror images) must be below 0.01.  The reference's B*P*P Python loop is a
    single nonzero() gather here.  Returns [B] (the global ratio, tiled)."""
    B, P, _ = trans.shape
    thre = 0.01
    rot_type, rot = rot.rot_type, rot.rot
    b, i, j = (contact_points[..., 0] == 1).nonzero(as_tuple=True)
    points1 = torch.stack(get_sym_point_list(contact_points[b, i, j, 1:]), 1)
    points2 = torch.stack(get_sym_point_list(contact_points[b, j, i, 1:]), 1)
    points1 = transform_pc(trans[b, i], rot[b, i], points1, rot_type=rot_type)
    points2 = transform_pc(trans[b, j], rot[b, j], points2, rot_type=rot_type)
    dist = ((points1[:, :, None] - points2[:, None, :])**2).sum(-1)
    dist = dist.flatten(1).min(-1)[0]
    acc = (dist < thre).sum().float() / float(dist.numel())
    return torch.ones(B).type_as(trans) * acc


@torch.no_grad()
def trans_metrics(trans1, trans2, valids, metric):
    """MSE / RMSE / MAE between translations (reference :145-168)."""
    assert metric in ['mse', 'rmse', 'mae']
    diff = trans1 - trans2
    if metric == 'mse':
        per_part = diff.pow(2).mean(dim=-1)
    elif metric == 'rmse':
        per_part = diff.pow(2).mean(dim=-1)**0.5
    else:
        per_part = diff.abs().mean(dim=-1)
    return _valid_mean(per_part, valids)


@torch.no_grad()
def rot_metrics(rot1, rot2, valids, metric):
    """MSE / RMSE / MAE between rotations in Euler degrees (reference :171-199)."""
    assert metric in ['mse', 'rmse', 'mae']
    delta = (rot1.to_euler(to_degree=True) - rot2.to_euler(to_degree=True)).abs()
    diff = torch.minimum(delta, 360. - delta)  # wrap-around at 180 deg
    if metric == 'mse':
        per_part = diff.pow(2).mean(dim=-1)
    elif metric == 'rmse':
        per_part = diff.pow(2).mean(dim=-1)**0.5
    else:
        per_part = diff.abs().mean(dim=-1)
    return _valid_mean(per_part, valids)
