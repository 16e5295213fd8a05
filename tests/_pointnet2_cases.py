"""Cases shared by tests/golden/make_golden_pointnet2.py (runs the REFERENCE's extensions/pointnet2/pointnet2_utils.py)
and tests/test_ref_pointnet2.py (replays them through this repo's pointnet2_utils)."""
import numpy as np
import torch

B, N, M, C = 2, 400, 24, 5

# QueryAndGroup constructor arguments per case (+ whether features are passed)
QAG = {
    "default": (dict(radius=0.25, nsample=16), True),
    "no_features": (dict(radius=0.3, nsample=8), False),
    "features_only": (dict(radius=0.25, nsample=16, use_xyz=False), True),
    "normalized_ret_xyz": (dict(radius=0.2, nsample=32, normalize_xyz=True, ret_grouped_xyz=True), True),
    "uniform_cnt": (dict(radius=0.15, nsample=16, sample_uniformly=True, ret_unique_cnt=True, ret_grouped_xyz=True), True),
}
GROUP_ALL = {"xyz_and_features": (dict(use_xyz=True), True), "features_only": (dict(use_xyz=False), True),
             "xyz_only": (dict(use_xyz=True), False)}


def inputs():
    from pointdae_b200 import synth
    xyz = synth.clouds(B, N, seed=77)
    rng = np.random.default_rng(77)
    new_xyz = np.ascontiguousarray(xyz[:, rng.permutation(N)[:M]])
    features = rng.standard_normal((B, C, N)).astype(np.float32)
    return xyz, new_xyz, features


def as_tuple(out):
    return out if isinstance(out, tuple) else (out,)


def weights_for(name, shape):
    """upstream gradient for output `name`, regenerated from the name"""
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    return torch.randn(shape, generator=g)
