"""Chamfer forward at the scene-scale configs (C4 256x8192^2 per 8 GPUs -> 32 clouds here, C5 1x100000^2): tensor-core filter in
column chunks against the FP32-pipe kernels.  python profiles/time_chamfer_big.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
L = _native.lib()
out = {}
for name, (bs, n) in {"C4 32x8192^2": (32, 8192), "C4 256x8192^2": (256, 8192), "C5 1x100000^2": (1, 100000), "16x4096^2": (16, 4096)}.items():
    c = synth.clouds(min(bs, 8), n, seed=5)
    reps = -(-bs // c.shape[0])
    import numpy as np
    c = np.tile(c, (reps, 1, 1))[:bs]
    a, b = torch.from_numpy(synth.prediction(c, seed=5)).to(dev), torch.from_numpy(c).to(dev)
    row = {}
    res = {}
    for mode in (0, 2, 3):
        L.pdae_tune_chamfer_tc(mode, -1.0)
        for _ in range(2):
            r = ops.chamfer_forward(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = ops.chamfer_forward(a, b)
        e1.record()
        torch.cuda.synchronize()
        row["mode%d_ms" % mode] = e0.elapsed_time(e1) / 5
        res[mode] = r
    row["same_as_fp32_pipe"] = all(torch.equal(x, y) for x, y in zip(res[0], res[3])) and all(torch.equal(x, y) for x, y in zip(res[0], res[2]))
    out[name] = row
    print(name, json.dumps(row), flush=True)
L.pdae_tune_chamfer_tc(3, -1.0)

# one rank's share of the reference-set-sharded C5 forward (world 8 and 2): all 100 000 rows against a slice
c = synth.clouds(1, 100000, seed=5)
a, b = torch.from_numpy(synth.prediction(c, seed=5)).to(dev), torch.from_numpy(c).to(dev)
for world in (8, 2):
    sl = b[:, : 100000 // world].contiguous()
    row = {}
    for mode in (0, 3):
        L.pdae_tune_chamfer_tc(mode, -1.0)
        for _ in range(2):
            ops.chamfer_sharded_local(a, sl, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.chamfer_sharded_local(a, sl, 0)
        e1.record()
        torch.cuda.synchronize()
        row["mode%d_ms" % mode] = e0.elapsed_time(e1) / 10
    print("sharded C5 share, world", world, json.dumps(row), flush=True)
L.pdae_tune_chamfer_tc(3, -1.0)
