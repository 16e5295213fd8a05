"""Cases shared by tests/golden/make_golden_group.py and tests/test_group_flavours.py."""
import numpy as np

# flavour -> (reference file under models/, input channels)
FLAVOURS = {
    "plain": ("PointCAE_transformer.py", 3),
    "with_index": ("Point_M2AE_modules.py", 3),
    "normal": ("MaskSurf.py", 6),
    "attribute": ("MaskSurf_v2.py", 7),
}
# case -> (B, N, num_group, group_size)
SHAPES = {"small": (2, 300, 12, 8), "ragged": (3, 1111, 20, 32), "dup": (2, 640, 16, 16)}


def inputs(case, b, n, channels):
    from pointdae_b200 import synth
    seed = sum(map(ord, case))
    xyz = synth.clouds(b, n, seed=seed)
    if case == "dup":
        xyz = synth.adversarial(xyz, seed=seed)  # near-origin points (FPS skip rule) and exact duplicates (kNN ties)
    if channels == 3:
        return xyz
    extra = np.random.default_rng(seed).standard_normal((b, n, channels - 3)).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([xyz, extra], axis=2))
