"""pointdae_b200 -- B200-native (sm_100a) implementation of Point-DAE's point-cloud geometry hot path.

Public surface mirrors the reference's names:
    pointnet2_utils.furthest_point_sample / gather_operation / ball_query / grouping_operation
    knn_cuda.KNN, group.Group, group.fps
    chamfer.forward / backward, chamfer_dist.ChamferDistanceL1 / L2 / L2_split / ChamferFunction
    dgcnn_util.knn / get_graph_feature
    install() / patch_models()   -- make the reference's own imports resolve here
    sharded.chamfer_forward_sharded -- reference-set sharding over torch.distributed
"""
from . import _native  # noqa: F401
from .install import install, patch_models  # noqa: F401

__all__ = ["install", "patch_models", "build"]


def build(verbose=False):
    return _native.build(verbose=verbose)
