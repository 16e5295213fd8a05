"""Drop-in for `pointnet2_ops.pointnet2_utils` (and the vendored `extensions/pointnet2/pointnet2_utils.py`)
: furthest_point_sample, gather_operation, three_nn, three_interpolate, grouping_operation, ball_query,
QueryAndGroup, GroupAll.

Same names, argument meaning and error behaviour as the reference
(extensions/pointnet2/pointnet2_utils.py:49-424); the compute is the sm_100a kernels.
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


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        # reference: pointnet2_utils.py:120-141 -> _ext.three_nn; Euclidean (not squared) distances out
        dist2, idx = ops.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        # reference: pointnet2_utils.py:152-176 -> _ext.three_interpolate
        m = features.size(2)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        return ops.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        # reference: pointnet2_utils.py:178-202: gradient to the features only
        idx, weight, m = ctx.three_interpolate_for_backward
        return ops.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = ThreeInterpolate.apply


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
    """reference: pointnet2_utils.py:287-374: radius grouping around new_xyz, neighbours re-centred (and
    optionally divided by the radius), features concatenated behind the xyz offsets."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz, self.normalize_xyz = ret_grouped_xyz, normalize_xyz
        self.sample_uniformly, self.ret_unique_cnt = sample_uniformly, ret_unique_cnt
        if ret_unique_cnt:
            assert sample_uniformly

    def _resample_uniformly(self, idx):
        """reference :333-343: every region keeps its distinct hits and fills the remaining slots by drawing
        among them (same torch.randint call sequence on the default CPU generator, so seeds reproduce)."""
        unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
        host = idx.cpu()
        for bi in range(host.shape[0]):
            for region in range(host.shape[1]):
                distinct = torch.unique(host[bi, region, :])
                unique_cnt[bi, region] = distinct.shape[0]
                fill = torch.randint(0, distinct.shape[0], (self.nsample - distinct.shape[0],), dtype=torch.long)
                host[bi, region, :] = torch.cat((distinct, distinct[fill]))
        idx.copy_(host)
        return unique_cnt

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = self._resample_uniformly(idx) if self.sample_uniformly else None
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B, 3, npoint, nsample)
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
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(torch.nn.Module):
    """reference: pointnet2_utils.py:377-424: the whole cloud as one group, (B, 3 + C, 1, N); pure reshaping.
    (The reference reads self.ret_grouped_xyz without ever setting it -- an AttributeError there; it is stored here.)"""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz, self.ret_grouped_xyz = use_xyz, ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            new_features = grouped_xyz
        elif self.use_xyz:
            new_features = torch.cat([grouped_xyz, features.unsqueeze(2)], dim=1)
        else:
            new_features = features.unsqueeze(2)
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features
