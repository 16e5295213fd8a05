"""Generates tests/golden/chamfer_losses.npz in THIS container (CPU only):

    python tests/golden/make_golden_losses.py

It imports the REFERENCE's own `extensions/chamfer_dist/__init__.py` from /root/reference (with `ipdb`
stubbed and the compiled `chamfer` module replaced by the oracle-backed stand-in tests/_oracle_chamfer.py,
which itself is pinned to the reference kernel's outputs by tests/golden/chamfer.npz), runs every loss
class the file exports on seeded inputs, and stores inputs, returned values and input gradients.
tests/test_chamfer_losses.py replays them through this repo's classes: on CPU over the same stand-in
(host arithmetic parity), on the GPU over the sm_100a kernels (end-to-end parity).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pointdae_b200 import synth  # noqa: E402
import _oracle_chamfer  # noqa: E402
import _loss_cases  # noqa: E402

REF = "/root/reference/extensions/chamfer_dist/__init__.py"


def load_reference():
    sys.modules["chamfer"] = _oracle_chamfer
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    spec = importlib.util.spec_from_file_location("ref_chamfer_dist", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    out = {}
    for name, (cls_name, arrays) in _loss_cases.cases(synth).items():
        results, grads = _loss_cases.run(getattr(ref, cls_name)(), arrays, torch.device("cpu"))
        for k, v in arrays.items():
            out["%s/in/%s" % (name, k)] = v
        for i, r in enumerate(results):
            out["%s/out/%d" % (name, i)] = r
        for k, g in grads.items():
            out["%s/grad/%s" % (name, k)] = g
        print(name, cls_name, [np.asarray(r).reshape(-1)[:1] for r in results])
    path = os.path.join(ROOT, "tests", "golden", "chamfer_losses.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
