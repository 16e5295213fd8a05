"""One tensor-core Chamfer forward at the headline shape per mode (for ncu): python profiles/run_chamfer_tc_once.py [mode ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
c = synth.clouds(128, 2048, seed=1)
a, b = torch.from_numpy(synth.prediction(c, seed=1)).to(dev), torch.from_numpy(c).to(dev)
for mode in [int(x) for x in sys.argv[1:]] or [1, 2]:
    _native.lib().pdae_tune_chamfer_tc(mode, -1.0)
    for _ in range(3):
        ops.chamfer_forward(a, b)
    torch.cuda.synchronize()
