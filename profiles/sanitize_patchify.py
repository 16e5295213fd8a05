"""Small single-launch patchifier calls for compute-sanitizer (memcheck / racecheck / synccheck): every (QW, NCW) shape,
ragged sizes, the corruption epilogue and the exact fallback (mass ties)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, ops, synth  # noqa: E402

L = _native.lib()
dev = torch.device("cuda:0")
for (b, n, g, m) in [(2, 1000, 9, 17), (2, 2048, 16, 32)]:
    xyz = synth.adversarial(synth.clouds(b, n, seed=n), seed=n)
    xyz[0, :300] = xyz[0, 0]  # a tie class larger than the candidate queue: streaming warp-select fallback
    X = torch.from_numpy(xyz).to(dev)
    mats = torch.from_numpy(np.random.default_rng(0).standard_normal((b, 2, 3, 3)).astype(np.float32))
    L.pdae_tune_patchify(0, 1, 8)
    want = ops.fps_group(X, g, m, want_idx=True)
    for qw in (1, 2):
        for ncw in (8, 12):
            L.pdae_tune_patchify(2, qw, ncw)
            got = ops.fps_group(X, g, m, want_idx=True)
            assert all(torch.equal(a, b_) for a, b_ in zip(got, want)), (qw, ncw)
            ops.fps_group_affine(X, g, m, mats)
torch.cuda.synchronize()
print("sanitize_patchify ok")
