"""Generates tests/golden/group_flavours.npz in THIS container (CPU only):

    python tests/golden/make_golden_group.py

The reference defines its Point-MAE-style patchifier `Group` four times with different outputs
(models/PointCAE_transformer.py:54-86, models/Point_M2AE_modules.py, models/MaskSurf.py, models/MaskSurf_v2.py).
The model files cannot be imported here (timm, compiled extensions), so this script lifts the `Group` class
statement -- and nothing else -- out of each file with `ast`, executes it UNMODIFIED in a namespace whose three
external names (`KNN`, `misc.fps`, `pointnet2_utils.gather_operation`) are oracle-backed stand-ins (oracle/cpu.py,
itself pinned to the reference kernels by tests/golden/fps_gather.npz), and stores what the reference's own
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
from oracle import cpu as oracle  # noqa: E402
import _group_cases as cases  # noqa: E402

REF = "/root/reference/models/"


class KNN(nn.Module):  # knn_cuda.KNN stand-in (transpose_mode=True is the only mode Group uses)
    def __init__(self, k, transpose_mode=False):
        super().__init__()
        assert transpose_mode
        self.k = k

    def forward(self, ref, query):
        d, i = oracle.knn(ref.numpy(), query.numpy(), self.k)
        return torch.from_numpy(d), torch.from_numpy(i)


def _gather_operation(features, idx):
    return torch.from_numpy(oracle.gather(features.numpy(), idx.numpy()))


def _fps(data, number):  # utils/misc.py:13-20
    fps_idx = torch.from_numpy(oracle.fps(data[:, :, :3].contiguous().numpy(), number))
    fps_data = _gather_operation(data.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
    return fps_idx, fps_data


def reference_group(filename):
    tree = ast.parse(open(REF + filename).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Group")
    ns = {"nn": nn, "torch": torch, "KNN": KNN, "misc": types.SimpleNamespace(fps=_fps),
          "pointnet2_utils": types.SimpleNamespace(gather_operation=_gather_operation)}
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
