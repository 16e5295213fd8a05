"""Column-split Chamfer forward: identical results for every chunk count, and CUDA-event timings per shape.
usage: python profiles/probe_split.py  -> one JSON object on stdout"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, ops, synth  # noqa: E402


def timed(fn, iters=60, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    L = _native.lib()
    dev = torch.device("cuda:0")
    out = {"unit": "us per forward (eager, CUDA events)"}
    shapes = [(128, 2048, 2048), (128, 1024, 1024), (16, 8192, 8192), (256, 8192, 8192), (1, 100000, 100000),
              (32, 2048, 2048), (8, 4096, 1536)]
    for b, n, m in shapes:
        a = torch.from_numpy(synth.clouds(min(b, 16), n, seed=n)).to(dev)
        c = torch.from_numpy(synth.clouds(min(b, 16), m, seed=m + 1)).to(dev)
        if b > 16:
            a = a.repeat((b + 15) // 16, 1, 1)[:b].contiguous()
            c = c.repeat((b + 15) // 16, 1, 1)[:b].contiguous()
            a += 1e-3 * torch.randn_like(a)
        L.pdae_tune_chamfer_split(1)
        want = ops.chamfer_forward(a, c)
        row = {}
        iters = 60 if b * n * m < 2e10 else 5
        for nc in (1, 2, 3, 4, 8, 0):
            L.pdae_tune_chamfer_split(nc)
            got = ops.chamfer_forward(a, c)
            same = all(torch.equal(x, y) for x, y in zip(got, want))
            row["auto" if nc == 0 else str(nc)] = [round(timed(lambda: ops.chamfer_forward(a, c), iters=iters, warm=3), 2), same]
        out["%dx%dx%d" % (b, n, m)] = row
    L.pdae_tune_chamfer_split(0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
