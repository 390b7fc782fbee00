from .encoder import *  # noqa: F401,F403
from .encoder import build_encoder, PointNet, DGCNN
from .regressor import PoseRegressor, StocasticPoseRegressor
from .base_model import BaseModel
