"""Drop-in for `pointnet2_ops.pointnet2_modules` (pointnet2_ops 3.0.0, un-vendored pip dependency of the reference,
README.md:43): the PointNet++ set-abstraction and feature-propagation modules the reference builds its PointNet++
encoder from -- `from pointnet2_ops.pointnet2_modules import PointnetFPModule, PointnetSAModule`
(models/pointnetv2_util.py:317-325: `PointnetSAModule(npoint=512, radius=0.2, nsample=32, mlp=[0, 64, 64, 128],
use_xyz=True)`, `sa(xyz, features) -> (new_xyz, new_features)`, `npoint=None` groups the whole cloud).
`models/__init__.py` imports that file, so without this module `import models` fails once `install()` has replaced
the `pointnet2_ops` package.

The package itself is not in the reference tree ("parity unpinned", like KNN_CUDA): constructor signatures, attribute
names (`groupers`, `mlps`, `mlp` -- they are the state-dict keys of the reference's checkpoints: `sa1.mlps.0.0.weight`
is the first 1x1 convolution, `.1` its BatchNorm) and the forward contract follow the published 3.0.0 API and are
anchored on the call sites above.  The geometry (FPS, gather, ball query, grouping, three_nn, interpolation) runs on
this repo's kernels through `pointnet2_utils`; the shared MLPs are ordinary torch layers."""
import torch
import torch.nn as nn

from . import pointnet2_utils


def build_shared_mlp(mlp_spec, bn=True):
    """[c0, c1, ..., cn] -> Sequential of (1x1 Conv2d, [BatchNorm2d], ReLU) per step; the convolution carries a bias
    only when no BatchNorm follows."""
    layers = []
    for c_in, c_out in zip(mlp_spec[:-1], mlp_spec[1:]):
        layers.append(nn.Conv2d(c_in, c_out, kernel_size=1, bias=not bn))
        if bn:
            layers.append(nn.BatchNorm2d(c_out))
        layers.append(nn.ReLU(True))
    return nn.Sequential(*layers)


class _PointnetSAModuleBase(nn.Module):
    """Sample `npoint` centres (FPS), group around each with every grouper, run the grouper's shared MLP and take the
    maximum over the group; the scales' features are concatenated."""

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz, features):
        """xyz (B,N,3), features (B,C,N) or None -> new_xyz (B,npoint,3) (None when the whole cloud is one group),
        new_features (B, sum_k mlps[k][-1], npoint)."""
        new_xyz = None
        if self.npoint is not None:
            picked = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
            new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), picked).transpose(1, 2).contiguous()
        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            grouped = mlp(grouper(xyz, new_xyz, features))           # (B, mlp[-1], npoint, nsample)
            pooled.append(torch.max(grouped, dim=3)[0])               # (B, mlp[-1], npoint)
        return new_xyz, torch.cat(pooled, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale grouping: one (radius, nsample, mlp) triple per scale."""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            if npoint is not None:
                self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz))
            else:
                self.groupers.append(pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3  # in place, like the package: the caller's list reflects the xyz channels afterwards
            self.mlps.append(build_shared_mlp(spec, bn))


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction; npoint=None (with radius / nsample None) pools the whole cloud."""

    def __init__(self, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetFPModule(nn.Module):
    """Feature propagation: inverse-distance interpolation of the three nearest known points' features, concatenated
    with the unknown points' own features, then a shared MLP."""

    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = build_shared_mlp(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        """unknown (B,n,3), known (B,m,3) or None, unknow_feats (B,C1,n) or None, known_feats (B,C2,m) -> (B,mlp[-1],n)."""
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            inverse = 1.0 / (dist + 1e-8)
            weight = inverse / torch.sum(inverse, dim=2, keepdim=True)
            spread = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            spread = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        stacked = spread if unknow_feats is None else torch.cat([spread, unknow_feats], dim=1)
        return self.mlp(stacked.unsqueeze(-1)).squeeze(-1)
