"""Generates tests/golden/group_flavours.npz in THIS container (CPU only):

    python tests/golden/make_golden_group.py

The reference defines its Point-MAE-style patchifier `Group` four times with different outputs
(models/PointCAE_transformer.py:54-86, models/Point_M2AE_modules.py, models/MaskSurf.py, models/MaskSurf_v2.py).
The model files cannot be imported here (timm, compiled extensions), so this script lifts the `Group` class
statement -- and nothing else -- out of each file with `ast`, executes it UNMODIFIED in a namespace whose three
external names (`KNN`, `misc.fps`, `pointnet2_utils.gather_operation`) are oracle-backed stand-ins
(tests/golden/_standins.py over oracle/cpu.py, itself pinned to the reference kernels by tests/golden/fps_gather.npz), and stores what the reference's own
forward returns on seeded inputs.  tests/test_group_flavours.py replays them: on CPU against the oracle
composition, on the GPU against this repo's fused classes."""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _group_cases as cases  # noqa: E402
import _standins  # noqa: E402

REF = "/root/reference/models/"


def reference_group(filename):
    tree = ast.parse(open(REF + filename).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Group")
    ns = {"nn": nn, "torch": torch, "KNN": _standins.KNN, "misc": _standins.misc,
          "pointnet2_utils": _standins.pointnet2_utils}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF + filename, "exec"), ns)
    return ns["Group"]


def main():
    out = {}
    for name, (filename, channels) in cases.FLAVOURS.items():
        cls = reference_group(filename)
        for case, (b, n, g, m) in cases.SHAPES.items():
            x = cases.inputs(case, b, n, channels)
            res = cls(g, m)(torch.from_numpy(x))
            for i, r in enumerate(res):
                out["%s/%s/%d" % (name, case, i)] = r.numpy()
            print(name, case, [tuple(r.shape) for r in res])
    path = os.path.join(ROOT, "tests", "golden", "group_flavours.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
