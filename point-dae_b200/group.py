"""Drop-in for the Point-MAE-style patchifier `Group(num_group, group_size)`
(models/PointCAE_transformer.py:54-86, byte-identical copies in models/Point_MAE.py:51-83 etc.)
and for `utils/misc.py:13-20 fps`.  Two launches (FPS+centre gather, kNN+gather+centre-subtract)
replace the reference's ~1500; clouds of 512..2048 points take ONE (csrc/patchify.cu)."""
import torch
import torch.nn as nn

from . import ops
from .knn_cuda import KNN


def _needs_grad(t):
    return torch.is_grad_enabled() and t.requires_grad


def fps(data, number):
    """utils/misc.py:13-20.  data (B,N,3|6) -> (fps_idx (B,G) int32, fps_data (B,G,3|6)).
    The reference gathers the centres with the differentiable `gather_operation`; every call site passes raw data, so
    the fused kernel (no autograd node) is the normal route, and a tensor that requires grad takes the differentiable
    gather through the same indices (same values, gradient scattered to the picked rows)."""
    fps_idx, fps_data = ops.fps_gather(data, number)
    if _needs_grad(data):
        fps_data = torch.gather(data, 1, fps_idx.long().unsqueeze(-1).expand(-1, -1, data.size(2)))
    return fps_idx, fps_data


class Group(nn.Module):  # FPS + KNN
    def __init__(self, num_group, group_size):
        super().__init__()
        self.num_group = num_group
        self.group_size = group_size
        self.knn = KNN(k=self.group_size, transpose_mode=True)  # kept for attribute parity

    def forward(self, xyz, knn_after=None):
        """input: B N 3  ->  neighborhood: B G M 3 (centre-subtracted), center: B G 3
        knn_after (optional, not in the reference): a torch.cuda.Event the kNN launch waits for on the current
        stream -- lets a caller that runs the patchifier on a side stream keep the issue-bound kNN out of the way of
        an FMA-bound kernel (ops.chamfer_scan_event); FPS is latency-bound and starts at once."""
        batch_size, num_points, _ = xyz.shape
        xyz = xyz.float().contiguous()
        if knn_after is None and not _needs_grad(xyz) and xyz.size(2) == 3:
            # one launch: the kNN of every centre runs beside the sampling of the next ones (ops.fps_group)
            _, center, neighborhood, _ = ops.fps_group(xyz.detach(), self.num_group, self.group_size)
            return neighborhood, center
        _, center = fps(xyz, self.num_group)  # B G 3
        if knn_after is not None:
            torch.cuda.current_stream().wait_event(knn_after)
        if _needs_grad(xyz):  # the reference's indexing and subtraction are differentiable (:80-85)
            _, idx = ops.group_points_knn(xyz.detach(), center.detach(), self.group_size, want_idx=True)
            return _gather_rows(xyz, idx) - center.unsqueeze(2), center
        neighborhood, _ = ops.group_points_knn(xyz.detach(), center, self.group_size, want_idx=False)
        return neighborhood, center

    def forward_corrupted(self, xyz, corrupt_type=('affine_r3',), mats=None):
        """The first seven lines of the reference model's forward in two launches
        (models/PointCAE_transformer.py:1010-1017: Group, `+ center`, `corrupt_data`, `- center` twice):
        -> neighborhood, center, transformed_neighborhood, transformed_center.
        `mats` (B,T,3,3) overrides the random draw of `corrupt_util_tensor.corrupt_stack(B, corrupt_type)`, which
        consumes the host RNGs exactly as the reference's `corrupt_data` does."""
        from . import corrupt_util_tensor
        xyz = xyz.float().contiguous()
        if mats is None:
            mats = corrupt_util_tensor.corrupt_stack(xyz.size(0), list(corrupt_type))
        if mats is None:  # clean / Drop-Patch only: the transformed copies are the clean ones, round trip included
            mats = torch.zeros((xyz.size(0), 0, 3, 3))
        if xyz.size(2) == 3:  # one call: the matrices come from the host RNGs and do not depend on the sampling
            _, center, nb, tnb, tc, _ = ops.fps_group_affine(xyz.detach(), self.num_group, self.group_size, mats)
            return nb, center, tnb, tc
        _, center = fps(xyz, self.num_group)
        nb, tnb, tc, _ = ops.group_affine(xyz.detach(), center, self.group_size, mats)
        return nb, center, tnb, tc


def _patches_with_idx(xyz_only, num_group, group_size):
    """FPS + centre gather, kNN + gather + centre-subtract with the int64 neighbour indices kept: two launches."""
    if not _needs_grad(xyz_only) and xyz_only.dim() == 3 and xyz_only.size(2) == 3:
        return ops.fps_group(xyz_only.detach().float().contiguous(), num_group, group_size, want_idx=True)
    fps_idx, center = fps(xyz_only, num_group)
    neighborhood, idx = ops.group_points_knn(xyz_only.detach(), center.detach(), group_size, want_idx=True)
    if _needs_grad(xyz_only):  # differentiable like the reference's indexing + subtraction
        neighborhood = _gather_rows(xyz_only, idx) - center.unsqueeze(2)
    return fps_idx, center, neighborhood, idx


def _gather_rows(attribute, idx):
    """attribute (B,N,A), idx (B,G,M) int64 -> (B,G,M,A): the reference's `view(B*N,-1)[idx + b*N]` (plain torch
    indexing, differentiable like the reference's)."""
    b, n, a = attribute.shape
    flat = (idx + torch.arange(b, device=idx.device).view(-1, 1, 1) * n).view(-1)
    return attribute.reshape(b * n, a)[flat, :].view(b, idx.size(1), idx.size(2), a).contiguous()


class GroupWithIndex(Group):
    """`Group` of models/Point_M2AE_modules.py:  -> neighborhood, center, idx, with idx the flattened batch-global
    neighbour indices `(idx + arange(B)*N).view(-1)` the multi-scale encoder reuses."""

    def forward(self, xyz):
        batch_size, num_points, _ = xyz.shape
        xyz = xyz.float().contiguous()
        _, center, neighborhood, idx = _patches_with_idx(xyz, self.num_group, self.group_size)
        idx = idx + torch.arange(0, batch_size, device=xyz.device).view(-1, 1, 1) * num_points
        return neighborhood, center, idx.view(-1)


class GroupNormal(Group):
    """`Group` of models/MaskSurf.py: input B N 6 (xyz + normal) -> neighborhood_no_normal (centre-subtracted),
    neighborhood_only_normal (the neighbours' normals, untouched), center."""

    def forward(self, xyz):
        xyz_no_normal = xyz[:, :, :3].float().contiguous()
        xyz_only_normal = xyz[:, :, 3:6].contiguous()
        _, center, neighborhood, idx = _patches_with_idx(xyz_no_normal, self.num_group, self.group_size)
        return neighborhood, _gather_rows(xyz_only_normal, idx), center


class GroupAttribute(Group):
    """`Group` of models/MaskSurf_v2.py, models/MaskFeat_transformer.py, models/MaskFeat_DGCNN.py: input B N 3+A ->
    neighborhood_xyz_only, neighborhood_attribute_only (B G M A), center, center_attribute (B G A)."""

    def forward(self, xyz):
        xyz_only = xyz[:, :, :3].float().contiguous()
        attribute_only = xyz[:, :, 3:].contiguous()
        fps_idx, center, neighborhood, idx = _patches_with_idx(xyz_only, self.num_group, self.group_size)
        b, n, a = attribute_only.shape
        flat = (fps_idx.long() + torch.arange(b, device=xyz.device).view(-1, 1) * n).view(-1)
        center_attribute = attribute_only.reshape(b * n, a)[flat, :].view(b, self.num_group, a).contiguous()
        return neighborhood, _gather_rows(attribute_only, idx), center, center_attribute
