"""Seeded cases shared by tests/golden/make_golden_corrupt.py (runs the REFERENCE's functions) and
tests/test_corrupt.py (replays them through this repo's host mirror, the oracle and the sm_100a kernels)."""
import random

import numpy as np
import torch

# name -> (reference function name, level argument, batch, groups, group size)
SINGLE = {
    "scale_nonorm_l4": ("corrupt_scale_nonorm", 4, 5, 7, 6),
    "scale_nonorm_l0": ("corrupt_scale_nonorm", 0, 3, 4, 9),
    "tranlate_l4": ("corrupt_tranlate", 4, 5, 7, 6),
    "rotate_l4": ("corrupt_rotate_360", 4, 6, 5, 8),
    "rotate_none": ("corrupt_rotate_360", None, 4, 5, 8),
    "rotate_z": ("corrupt_rotate_z_360", None, 4, 3, 5),
    "reflection": ("corrupt_reflection", None, 9, 4, 4),
    "shear_l4": ("corrupt_shear", 4, 5, 6, 7),
    "shear_none": ("corrupt_shear", None, 3, 2, 33),
}
# corrupt_data(type=...) cases: name -> (type list, batch, groups, group size)
CHAINS = {
    "affine_r3_s%d" % s: (["affine_r3"], 4 + s % 3, 8, 5) for s in range(12)
}
CHAINS["clean"] = (["clean"], 3, 4, 4)
CHAINS["droppatch_affine"] = (["Drop-Patch", "affine_r3"], 3, 64, 32)
CHAINS["affine_twice"] = (["affine_r3", "affine_r3"], 2, 6, 3)


def seed_all(name):
    s = sum(ord(ch) * (i + 1) for i, ch in enumerate(name)) % 100003
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)
    return s


def inputs(name, b, g, m):
    """Absolute-coordinate patches (B,G,M,3) and centres (B,G,3), drawn from their own generator so the corruption's
    RNG stream starts right after seed_all(name)."""
    rng = np.random.default_rng(sum(map(ord, name)))
    center = rng.uniform(-1, 1, size=(b, g, 3)).astype(np.float32)
    nb = (center[:, :, None, :] + 0.1 * rng.standard_normal((b, g, m, 3))).astype(np.float32)
    nb[0, 0, 0] = 0.0  # exact zeros survive every affine map
    return nb, center


def next_draws():
    """One draw from each host generator the corruptions use: equal values <=> equal stream positions."""
    return np.array([random.random(), np.random.rand(), float(torch.rand(1, dtype=torch.float64))], dtype=np.float64)


# dropout_patch_random(pc, level): name -> (batch, points, level)
DROP_PATCH = {"drop_level_none": (3, 1024, None), "drop_level_0": (2, 700, 0), "drop_level_4": (2, 2048, 4),
              "drop_all_masked": (1, 256, 60)}  # level 60 -> prob 6.5: no patch survives the draw, patch 0 is forced


def drop_patch_input(name, b, n):
    from pointdae_b200 import synth
    return synth.adversarial(synth.clouds(b, n, seed=sum(map(ord, name))), seed=1)
