"""Drop-in for `pointnet2_ops.pointnet2_utils` (and the vendored `extensions/pointnet2/pointnet2_utils.py`):
furthest_point_sample, gather_operation, three_nn, three_interpolate, grouping_operation, ball_query, QueryAndGroup,
GroupAll -- same public names, argument order, return values and error behaviour as the reference
(extensions/pointnet2/pointnet2_utils.py:49-424); every computation is an sm_100a kernel behind `ops`.

The autograd glue is written once (`_op`): each public op is a torch.autograd.Function whose forward calls the kernel
wrapper and whose backward, where the reference defines one, calls the matching gradient kernel.
"""
import torch
from torch.autograd import Function

from . import ops


def _op(name, forward, backward=None, index_outputs=(), n_inputs=2, doc=""):
    """Builds the Function class `name`.  forward(ctx, *inputs) -> output(s); backward(ctx, *grads) -> grad of the FIRST
    input (every other input is an index / weight / scalar without gradient in the reference as well).
    index_outputs: positions of integer outputs, marked non-differentiable."""

    def _forward(ctx, *inputs):
        out = forward(ctx, *inputs)
        outs = out if isinstance(out, tuple) else (out,)
        for pos in index_outputs:
            ctx.mark_non_differentiable(outs[pos])
        return out

    def _backward(ctx, *grads):
        first = backward(ctx, *grads) if backward is not None else None
        return (first,) + (None,) * (n_inputs - 1)

    return type(name, (Function,), {"forward": staticmethod(_forward), "backward": staticmethod(_backward), "__doc__": doc})


# ---- sampling ------------------------------------------------------------------------------------------------------
FurthestPointSampling = _op(
    "FurthestPointSampling", lambda ctx, xyz, npoint: ops.furthest_point_sample(xyz, npoint), index_outputs=(0,),
    doc="xyz (B,N,3) f32, npoint -> (B,npoint) int32 indices; reference :51-78 -> _ext.furthest_point_sampling")
furthest_point_sample = FurthestPointSampling.apply


def _gather_fwd(ctx, feats, index):
    ctx.gather_state = (index, feats.size(2))
    return ops.gather_points(feats, index)


def _gather_bwd(ctx, grad):
    index, n = ctx.gather_state
    return ops.gather_points_grad(grad.contiguous(), index, n)


GatherOperation = _op("GatherOperation", _gather_fwd, _gather_bwd,
                      doc="features (B,C,N), idx (B,M) int32 -> (B,C,M); reference :83-115 -> _ext.gather_points(_grad)")
gather_operation = GatherOperation.apply


# ---- feature propagation ---------------------------------------------------------------------------------------------
def _three_nn_fwd(ctx, unknown, known):
    squared, index = ops.three_nn(unknown, known)
    return torch.sqrt(squared), index  # the reference returns Euclidean distances (:139-141)


ThreeNN = _op("ThreeNN", _three_nn_fwd, index_outputs=(1,),
              doc="unknown (B,n,3), known (B,m,3) -> (dist (B,n,3), idx (B,n,3) int32) of the 3 nearest; reference :120-146")
three_nn = ThreeNN.apply


def _interp_fwd(ctx, feats, index, weight):
    ctx.interp_state = (index, weight, feats.size(2))
    return ops.three_interpolate(feats, index, weight)


def _interp_bwd(ctx, grad):
    index, weight, m = ctx.interp_state
    return ops.three_interpolate_grad(grad.contiguous(), index, weight, m)


ThreeInterpolate = _op("ThreeInterpolate", _interp_fwd, _interp_bwd, n_inputs=3,
                       doc="features (B,c,m), idx / weight (B,n,3) -> (B,c,n); gradient to the features only; reference :152-202")
three_interpolate = ThreeInterpolate.apply


# ---- grouping ----------------------------------------------------------------------------------------------------------
def _group_fwd(ctx, feats, index):
    ctx.group_state = (index, feats.size(2))
    return ops.group_points(feats, index)


def _group_bwd(ctx, grad):
    index, n = ctx.group_state
    return ops.group_points_grad(grad.contiguous(), index, n)


GroupingOperation = _op("GroupingOperation", _group_fwd, _group_bwd,
                        doc="features (B,C,N), idx (B,npoint,nsample) int32 -> (B,C,npoint,nsample); reference :260-300")
grouping_operation = GroupingOperation.apply

BallQuery = _op("BallQuery", lambda ctx, radius, nsample, xyz, new_xyz: ops.ball_query(new_xyz, xyz, radius, nsample),
                index_outputs=(0,), n_inputs=4,
                doc="radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3) -> (B,npoint,nsample) int32; reference :312-340")
ball_query = BallQuery.apply


class QueryAndGroup(torch.nn.Module):
    """reference :287-374: radius grouping around `new_xyz`; the neighbours' coordinates are re-centred (optionally
    divided by the radius) and put in front of their features."""

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
        counts = torch.zeros((idx.shape[0], idx.shape[1]))
        host = idx.cpu()
        for bi in range(host.shape[0]):
            for region in range(host.shape[1]):
                distinct = torch.unique(host[bi, region, :])
                counts[bi, region] = distinct.shape[0]
                fill = torch.randint(0, distinct.shape[0], (self.nsample - distinct.shape[0],), dtype=torch.long)
                host[bi, region, :] = torch.cat((distinct, distinct[fill]))
        idx.copy_(host)
        return counts

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = self._resample_uniformly(idx) if self.sample_uniformly else None
        offsets = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            offsets = offsets / self.radius
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            grouped = offsets  # (B, 3, npoint, nsample)
        else:
            picked = grouping_operation(features, idx)
            grouped = torch.cat([offsets, picked], dim=1) if self.use_xyz else picked  # (B, 3 + C, npoint, nsample)
        extras = ([offsets] if self.ret_grouped_xyz else []) + ([unique_cnt] if self.ret_unique_cnt else [])
        return tuple([grouped] + extras) if extras else grouped


class GroupAll(torch.nn.Module):
    """reference :377-424: the whole cloud as one group, (B, 3 + C, 1, N); pure reshaping.
    (The reference reads self.ret_grouped_xyz without ever setting it -- an AttributeError there; it is stored here.)"""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz, self.ret_grouped_xyz = use_xyz, ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        coords = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            whole = coords
        else:
            whole = torch.cat([coords, features.unsqueeze(2)], dim=1) if self.use_xyz else features.unsqueeze(2)
        return (whole, coords) if self.ret_grouped_xyz else whole
