"""Generates tests/golden/forward_seq.npz in THIS container (CPU only):

    python tests/golden/make_golden_forward_seq.py

The first seven lines of the reference model's forward (models/PointCAE_transformer.py:1010-1017) executed with the
reference's OWN pieces: its `Group` class statement (lifted with ast, as make_golden_group.py does) over the
oracle-backed stand-ins, its own `corrupt_data` (datasets/corrupt_util_tensor.py:706-728) on seeded host RNGs, and
the `+ center` / `- center` arithmetic spelled exactly as the model spells it.  Stored: neighborhood, center,
transformed_neighborhood, transformed_center.  `Group.forward_corrupted` must reproduce them from the same seeds."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
import _corrupt_cases as ccases  # noqa: E402
import make_golden_corrupt  # noqa: E402
import make_golden_group  # noqa: E402

# case -> (B, N, num_group, group_size, corrupt_type)
CASES = {"affine_a": (3, 700, 16, 8, ["affine_r3"]), "affine_b": (2, 1024, 64, 32, ["Drop-Patch", "affine_r3"]),
         "affine_c": (4, 300, 8, 16, ["affine_r3"]), "clean": (2, 256, 8, 8, ["clean"])}


def cloud(name, b, n):
    from pointdae_b200 import synth
    return synth.adversarial(synth.clouds(b, n, seed=sum(map(ord, name))), seed=3)


def main():
    ref_corrupt = make_golden_corrupt.load_reference()
    group_cls = make_golden_group.reference_group("PointCAE_transformer.py")
    out = {}
    for name, (b, n, g, m, typ) in CASES.items():
        pts = torch.from_numpy(cloud(name, b, n))
        ccases.seed_all(name)
        # models/PointCAE_transformer.py:1010-1017, verbatim in behaviour
        pts = pts[:, :, :3].contiguous()
        neighborhood, center = group_cls(g, m)(pts)
        neighborhood = neighborhood + center.unsqueeze(2)
        transformed_neighborhood, transformed_center = ref_corrupt.corrupt_data(neighborhood, center, type=typ)
        neighborhood = neighborhood - center.unsqueeze(2)
        transformed_neighborhood = transformed_neighborhood - transformed_center.unsqueeze(2)
        for key, t in (("neighborhood", neighborhood), ("center", center), ("t_neighborhood", transformed_neighborhood),
                       ("t_center", transformed_center)):
            out["%s/%s" % (name, key)] = t.numpy()
        out[name + "/rng_after"] = ccases.next_draws()
        print(name, typ, tuple(transformed_neighborhood.shape))
    path = os.path.join(ROOT, "tests", "golden", "forward_seq.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
