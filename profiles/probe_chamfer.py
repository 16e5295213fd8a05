"""Launches the Chamfer forward once per (variant, shape) so that `ncu --metrics ...` can attribute FMA-pipe
utilisation to the in-loop schedule vs the wave tail.  usage: python profiles/probe_chamfer.py v:B:N [v:B:N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pointdae_b200 import _native, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
L = _native.lib()
for spec in sys.argv[1:]:
    v, b, n = (int(x) for x in spec.split(":"))
    base = torch.from_numpy(synth.clouds(min(b, 8), n, seed=7)).to(dev)
    c = base.repeat((b + base.size(0) - 1) // base.size(0), 1, 1)[:b].contiguous()
    p = (c + 0.02 * torch.randn_like(c)).contiguous()
    L.pdae_tune_chamfer_variant(v)
    for _ in range(3):
        ops.chamfer_forward(p, c)
    torch.cuda.synchronize()
L.pdae_tune_chamfer_variant(0)
