"""Drop-in for the functions of `datasets/corrupt_util_tensor.py` that run inside the model's forward, next to the
patchifier:

* `dropout_patch_random` (:592-616), the Drop-Patch corruption: FPS to 64 centres, KNN 32, gather of the patches,
  random subset of the patches.  Here the patchifier is two launches (FPS + centre gather, kNN + patch gather fused).
* the affine family `corrupt_scale_nonorm` (:59-85), `corrupt_tranlate` (:88-113; it multiplies, like the reference),
  `corrupt_rotate_360` (:139-193), `corrupt_rotate_z_360` (:195-248), `corrupt_reflection` (:251-290),
  `corrupt_shear` (:306-343) and `corrupt_data` (:706-728), which chains one to three of them (`affine_r3`).  The
  reference launches one product / batched matmul per corruption and tensor; here the per-cloud 3x3 matrices are
  collected on the host and ONE kernel (`pdae_affine_points_f32`) applies the chain to patches and centres.
  `Group.forward_corrupted` (group.py) goes one step further and emits the corrupted patches from the kNN epilogue.

Random numbers are drawn exactly as in the reference -- same generators (`random`, `numpy.random`, torch's CPU
generator), same calls, same order -- so the same seeds produce the same matrices and the same patch masks
(tests/golden/corrupt.npz, made by running the reference's own functions).  The additive / dropout corruptions of that
file (`jitter`, `add_*`, `dropout_*`, `scan`) cannot be reached through the reference's `corrupt_data` (its generic
branch reads an unbound `level`, :722) and stay out."""
import math
import random

import numpy as np
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


# ------------------------------------------------------------------------------------------ matrices (host, CPU RNG)
def _eye(batch_size):
    return torch.eye(3).expand((batch_size, 3, 3)).clone().float()


def scale_nonorm_matrix(batch_size, level):
    s = [1.6, 1.7, 1.8, 1.9, 2.0][level]
    v = torch.FloatTensor(batch_size, 1, 1, 3).uniform_(1. / s, s)
    return torch.diag_embed(v.view(batch_size, 3))


def tranlate_matrix(batch_size, level):
    s = [0.1, 0.2, 0.3, 0.4, 0.5][level]
    v = torch.FloatTensor(batch_size, 1, 1, 3).uniform_(-s, s)
    return torch.diag_embed(v.view(batch_size, 3))  # the reference multiplies by the offsets (:111-113)


def _rotation(angles, axes):
    batch_size = angles.size(0)
    R = None
    for axis in axes:  # x first, z last: R = Rz @ (Ry @ Rx)
        c, s = torch.cos(angles[:, axis]), torch.sin(angles[:, axis])
        M = _eye(batch_size)
        i, j = [(1, 2), (2, 0), (0, 1)][axis]  # the plane the axis leaves fixed, oriented as in the reference
        M[:, i, i], M[:, i, j], M[:, j, i], M[:, j, j] = c, -s, s, c
        R = M if R is None else torch.matmul(M, R)
    return R


def rotate_360_matrix(batch_size, level=None):
    if level == None:  # noqa: E711
        level = random.random() * 4
    angle_clip = math.pi
    angle_clip = angle_clip / 5 * (level + 1)
    angles = torch.FloatTensor(batch_size, 3).uniform_(-angle_clip, angle_clip)
    return _rotation(angles, (0, 1, 2))


def rotate_z_360_matrix(batch_size, level=None):
    if level == None:  # noqa: E711  (drawn and unused, as in the reference :209-210)
        level = random.random() * 4
    angle_clip = math.pi
    angles = torch.FloatTensor(batch_size, 3).uniform_(-angle_clip, angle_clip)
    return _rotation(angles, (2,))


def reflection_matrix(batch_size, level=None):
    reflection = torch.from_numpy(np.random.choice(np.array([1, -1]), size=(batch_size, 3)))
    R = _eye(batch_size)
    # the reference writes the third sign into element [0][0] of its "Rz" as well (:278): diag(r0*r2, r1, 1)
    R[:, 0, 0] = (reflection[:, 0] * reflection[:, 2]).float()
    R[:, 1, 1] = reflection[:, 1].float()
    return R


def shear_matrix(batch_size, level=None):
    if level == None:  # noqa: E711
        level = random.random() * 4
    shear_clip = (level + 1) * 0.1
    shear = torch.from_numpy(np.random.uniform(-shear_clip, shear_clip, size=(batch_size, 6)))
    R = _eye(batch_size)
    R[:, 0, 1], R[:, 0, 2] = shear[:, 0], shear[:, 1]
    R[:, 1, 0], R[:, 1, 2] = shear[:, 2], shear[:, 3]
    R[:, 2, 0], R[:, 2, 1] = shear[:, 4], shear[:, 5]
    return R


affine_matrices = {
    'translate': tranlate_matrix,
    'scale_nonorm': scale_nonorm_matrix,
    'rotate': rotate_360_matrix,
    'rotate_z': rotate_z_360_matrix,
    'reflection': reflection_matrix,
    'shear': shear_matrix,
}
affine_corruptions = ['translate', 'scale_nonorm', 'rotate', 'reflection', 'shear']  # the pool of `affine_r3` (:702)


def corrupt_stack(batch_size, type=['clean']):
    """The (B,T,3,3) CPU tensor of matrices `corrupt_data(..., type)` would apply, in order (None when it applies
    nothing).  Consumes the host RNGs exactly as the reference's `corrupt_data` does."""
    mats = []
    level_bound = False  # the reference's `level = 4` (:718) stays bound for later items of the same call
    for corruption_item in type:
        if corruption_item == 'clean' or corruption_item == 'Drop-Patch':
            pass
        elif corruption_item == 'affine_r3':
            number = random.choice([1, 2, 3])
            for name in random.sample(affine_corruptions, number):
                mats.append(affine_matrices[name](batch_size, 4))
                level_bound = True
        elif level_bound and corruption_item in affine_matrices:
            # generic branch (:719-723) after an 'affine_r3' item, e.g. type=['affine_r3', 'rotate_z']: level is 4
            mats.append(affine_matrices[corruption_item](batch_size, 4))
        elif level_bound:
            raise KeyError(corruption_item)  # corruptions[corruption_item] (:723): not an affine corruption of this module
        else:
            # the reference's generic branch evaluates an unbound `level` (:723): same outcome, stated plainly
            raise NameError("corrupt_data: corruption %r needs a `level` the reference never defines "
                            "(datasets/corrupt_util_tensor.py:723)" % (corruption_item,))
    return torch.stack(mats, dim=1) if mats else None


# ------------------------------------------------------------------------------------------ application (one launch)
def _batch_of(pointcloud):
    first = pointcloud[0] if isinstance(pointcloud, list) else pointcloud
    return first.size(0), first.device


def _apply(pointcloud, center, mats):
    if isinstance(pointcloud, list):  # multi-scale models pass lists (models/Point_M2AE.py:799)
        out = [ops.affine_points(p, c, mats) for p, c in zip(pointcloud, center)]
        return [o[0] for o in out], [o[1] for o in out]
    return ops.affine_points(pointcloud, center, mats)


def _single(matrix_fn):
    def corrupt(pointcloud, center, level=None):
        batch_size, _ = _batch_of(pointcloud)
        return _apply(pointcloud, center, matrix_fn(batch_size, level).unsqueeze(1))
    return corrupt


corrupt_scale_nonorm = _single(scale_nonorm_matrix)
corrupt_tranlate = _single(tranlate_matrix)
corrupt_rotate_360 = _single(rotate_360_matrix)
corrupt_rotate_z_360 = _single(rotate_z_360_matrix)
corrupt_reflection = _single(reflection_matrix)
corrupt_shear = _single(shear_matrix)

corruptions = {
    'translate': corrupt_tranlate,
    'scale_nonorm': corrupt_scale_nonorm,
    'rotate': corrupt_rotate_360,
    'rotate_z': corrupt_rotate_z_360,
    'reflection': corrupt_reflection,
    'shear': corrupt_shear,
}


def corrupt_data(neighborhood, center, type=['clean']):
    """neighborhood (B,G,M,3) in absolute coordinates, center (B,G,3) (or lists of them) -> the corrupted pair."""
    batch_size, _ = _batch_of(neighborhood)
    mats = corrupt_stack(batch_size, type)
    if mats is None:
        return neighborhood, center
    return _apply(neighborhood, center, mats)
