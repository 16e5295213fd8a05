"""Drop-in for the Point-MAE-style patchifier `Group(num_group, group_size)`
(models/PointCAE_transformer.py:54-86, byte-identical copies in models/Point_MAE.py:51-83 etc.)
and for `utils/misc.py:13-20 fps`.  Two launches (FPS+centre gather, kNN+gather+centre-subtract)
replace the reference's ~1500."""
import torch
import torch.nn as nn

from . import ops
from .knn_cuda import KNN


def fps(data, number):
    """utils/misc.py:13-20.  data (B,N,3|6) -> (fps_idx (B,G) int32, fps_data (B,G,3|6))."""
    return ops.fps_gather(data, number)


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
        _, center = fps(xyz, self.num_group)  # B G 3
        if knn_after is not None:
            torch.cuda.current_stream().wait_event(knn_after)
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
        _, center = fps(xyz, self.num_group)
        if mats is None:
            mats = corrupt_util_tensor.corrupt_stack(xyz.size(0), list(corrupt_type))
        if mats is None:  # clean / Drop-Patch only: the transformed copies are the clean ones, round trip included
            mats = torch.zeros((xyz.size(0), 0, 3, 3))
        nb, tnb, tc, _ = ops.group_affine(xyz.detach(), center, self.group_size, mats)
        return nb, center, tnb, tc
