from .chamfer import chamfer_distance, nn_distance, chamfer_forward, \
    chamfer_backward, ChamferDistanceFunction
