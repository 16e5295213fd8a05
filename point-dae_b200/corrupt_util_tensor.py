"""Drop-in for the one function of `datasets/corrupt_util_tensor.py` that sits on the geometry hot path:
`dropout_patch_random` (:592-616), the Drop-Patch corruption executed inside the model's forward -- FPS to 64
centres, KNN 32, gather of the patches, random subset of the patches.  The rest of that file (affine / jitter /
density corruptions) is elementwise torch code with host-side RNG and keeps running from the reference unchanged.

Here the patchifier is two launches (FPS + centre gather, kNN + patch gather fused); the random numbers are drawn
exactly as in the reference (`random.random()` for the level, `torch.rand(64)` on the CPU generator for the mask), so
the same seeds select the same patches."""
import random

import torch

from . import ops

NUM_GROUP, GROUP_SIZE = 64, 32  # hard-coded in the reference (:597, :591, :601-602)


def dropout_patch_random(pc_tensor, level=None):
    """pc_tensor (B,N,3) -> (B, kept_groups*32, 3): the points of a random subset of the 64 FPS/KNN patches."""
    if level == None:  # noqa: E711  (same test as the reference: level 0 is a valid argument)
        level = random.random() * 4
    prob = level / 10.0 + 0.5
    batch_size = pc_tensor.shape[0]
    xyz = pc_tensor[:, :, :3].contiguous()
    _, center = ops.fps_gather(xyz, NUM_GROUP)
    patches, _ = ops.group_points_knn(xyz, center, GROUP_SIZE, want_idx=False, subtract_center=False)  # B 64 32 3
    group_mask = torch.rand(NUM_GROUP) > prob
    if group_mask.sum().item() == 0:  # at least one patch survives
        group_mask[0] = True
    return patches[:, group_mask.to(pc_tensor.device)].view(batch_size, -1, 3)
