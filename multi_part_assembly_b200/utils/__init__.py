"""Same re-exports as the reference's multi_part_assembly/utils/__init__.py:1-12
(minus the wandb/ply/pickle helpers, which are outside the hot path)."""
from .transforms import *  # noqa: F401,F403
from .transforms import random_quaternions, qmul, qrmat, qrot, qtransform, \
    qtransform_invert, rmat_rot, rmat_transform, rot_pc, transform_pc
from .rotation import Rotation3D, rot6d_to_matrix
from .chamfer import chamfer_distance
from .loss import trans_l2_loss, rot_l2_loss, rot_cosine_loss, \
    rot_points_l2_loss, rot_points_cd_loss, shape_cd_loss, repulsion_cd_loss
from .utils import filter_wd_parameters, _get_clones
from .eval_utils import trans_metrics, rot_metrics, calc_part_acc, \
    calc_connectivity_acc
from .config_utils import merge_cfg
