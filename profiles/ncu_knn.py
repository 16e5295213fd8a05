"""One launch per BASELINE shape of the Group kernel, for ncu (profiles/r02/call_*ncu*.sh):
    ncu --set full -k regex:knn4 ... python profiles/ncu_knn.py H C4 C5 C3"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointdae_b200 import group, ops, synth  # noqa: E402

SHAPES = {"H": (128, 2048, 64, 32, False), "C2": (128, 1024, 64, 32, False), "C4": (256, 8192, 512, 32, False),
          "C5": (1, 100000, 2048, 64, False), "C3": (16, 2048, 2048, 20, True)}
for name in sys.argv[1:] or ["H"]:
    b, r, q, k, planar = SHAPES[name]
    xyz = torch.from_numpy(synth.clouds(b, r, seed=7 + r)).to("cuda:0")
    if planar:
        x = xyz.transpose(1, 2).contiguous()
        for _ in range(2):
            ops.feat_knn(x, k)
    else:
        _, center = group.fps(xyz, q)
        center = center.contiguous()
        for _ in range(2):
            ops.group_points_knn(xyz, center, k, want_idx=True)
    torch.cuda.synchronize()
