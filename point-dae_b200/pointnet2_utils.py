"""Drop-in for `pointnet2_ops.pointnet2_utils` (and the vendored `extensions/pointnet2/pointnet2_utils.py`)
restricted to the hot path: furthest_point_sample, gather_operation, ball_query, grouping_operation.

Same names, argument meaning and error behaviour as the reference
(extensions/pointnet2/pointnet2_utils.py:49-115, 258-345); the compute is the sm_100a kernels.
"""
import torch
from torch.autograd import Function

from . import ops


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        # reference: pointnet2_utils.py:51-71 -> _ext.furthest_point_sampling
        fps_inds = ops.furthest_point_sample(xyz, npoint)
        ctx.mark_non_differentiable(fps_inds)
        return fps_inds

    @staticmethod
    def backward(xyz, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        # reference: pointnet2_utils.py:83-104 -> _ext.gather_points
        _, C, N = features.size()
        ctx.for_backwards = (idx, C, N)
        return ops.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        grad_features = ops.gather_points_grad(grad_out.contiguous(), idx, N)
        return grad_features, None


gather_operation = GatherOperation.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        # reference: pointnet2_utils.py:260-283 -> _ext.group_points
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        ctx.for_backwards = (idx, N)
        return ops.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        grad_features = ops.group_points_grad(grad_out.contiguous(), idx, N)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        # reference: pointnet2_utils.py:312-337 -> _ext.ball_query(new_xyz, xyz, radius, nsample)
        inds = ops.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(torch.nn.Module):
    """reference: pointnet2_utils.py:348-424 (radius grouping + optional xyz concat), the subset of
    options the 3DETR config uses (use_xyz, normalize_xyz, ret_grouped_xyz, ret_unique_cnt=False)."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        if sample_uniformly or ret_unique_cnt:
            raise NotImplementedError("sample_uniformly / ret_unique_cnt are outside the hot path")
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz, self.normalize_xyz = ret_grouped_xyz, normalize_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        return ret[0] if len(ret) == 1 else tuple(ret)
