"""Generates tests/golden/corrupt.npz in THIS container (CPU only):

    python tests/golden/make_golden_corrupt.py

It imports the REFERENCE's own `datasets/corrupt_util_tensor.py` from /root/reference (with `ipdb`, `knn_cuda` and
`pointnet2_ops` served by the oracle-backed stand-ins of tests/golden/_standins.py), seeds `random`,
`numpy.random` and torch's CPU generator, and runs the reference's affine corruptions and `corrupt_data` on CPU
tensors.  Inputs are regenerated from tests/_corrupt_cases.py; outputs are stored.  tests/test_corrupt.py replays
the same seeds through this repo's host mirror + oracle (CPU) and + sm_100a kernels (GPU).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _corrupt_cases as cases  # noqa: E402
import _standins  # noqa: E402

REF = "/root/reference/datasets/corrupt_util_tensor.py"


def load_reference():
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    _standins.install_modules()  # knn_cuda / pointnet2_ops: oracle-backed (dropout_patch_random runs on them)
    spec = importlib.util.spec_from_file_location("ref_corrupt_util_tensor", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    out = {}
    for name, (fn, level, b, g, m) in cases.SINGLE.items():
        nb, c = cases.inputs(name, b, g, m)
        cases.seed_all(name)
        tn, tc = getattr(ref, fn)(torch.from_numpy(nb), torch.from_numpy(c), level)
        out["single/%s/points" % name], out["single/%s/center" % name] = tn.numpy(), tc.numpy()
        out["single/%s/rng_after" % name] = cases.next_draws()  # pins how much of each host RNG stream was consumed
        print(name, fn, float(np.abs(tn.numpy()).max()))
    for name, (typ, b, g, m) in cases.CHAINS.items():
        nb, c = cases.inputs(name, b, g, m)
        cases.seed_all(name)
        tn, tc = ref.corrupt_data(torch.from_numpy(nb), torch.from_numpy(c), type=typ)
        out["chain/%s/points" % name], out["chain/%s/center" % name] = tn.numpy(), tc.numpy()
        out["chain/%s/rng_after" % name] = cases.next_draws()
        # the list form used by the multi-scale model (models/Point_M2AE.py:799)
        cases.seed_all(name)
        tl, cl = ref.corrupt_data([torch.from_numpy(nb), torch.from_numpy(nb[:, :2])],
                                  [torch.from_numpy(c), torch.from_numpy(c[:, :2])], type=typ)
        out["chain/%s/list1_points" % name], out["chain/%s/list1_center" % name] = tl[1].numpy(), cl[1].numpy()
        print(name, typ, float(np.abs(tn.numpy()).max()))
    # Drop-Patch (:592-616): FPS 64 + KNN 32 + patch gather + random patch subset, the reference's own function
    for name, (b, n, level) in cases.DROP_PATCH.items():
        pc = cases.drop_patch_input(name, b, n)
        cases.seed_all(name)
        kept = ref.dropout_patch_random(torch.from_numpy(pc), level)
        out["drop_patch/%s/points" % name] = kept.numpy()
        out["drop_patch/%s/rng_after" % name] = cases.next_draws()
        print(name, "dropout_patch_random", tuple(kept.shape))
    path = os.path.join(ROOT, "tests", "golden", "corrupt.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
