"""Timing of the single-launch patchifier (csrc/patchify.cu) against the two-launch form, every task width / consumer
count, CUDA-graph replays (4 calls per graph) timed with CUDA events.  Writes gpurun_out/time_patchify.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
L = _native.lib()


def timed_us(fn, reps=40):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [fn() for _ in range(4)]
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    del keep
    return a.elapsed_time(b) * 1e3 / (4 * reps)


out = {}
for (B, N, G, M) in [(128, 2048, 64, 32), (128, 1024, 64, 32), (16, 2048, 128, 32), (148, 2048, 64, 32), (256, 2048, 64, 32)]:
    c = torch.from_numpy(synth.clouds(B, N, seed=N)).to(dev)
    rec = {}
    L.pdae_tune_patchify(0, 2, 12)
    rec["fps_gather"] = timed_us(lambda: ops.fps_gather(c, G))
    center = ops.fps_gather(c, G)[1]
    rec["group_knn"] = timed_us(lambda: ops.group_points_knn(c, center, M, want_idx=False))
    rec["two_launch"] = timed_us(lambda: ops.fps_group(c, G, M))
    want = ops.fps_group(c, G, M, want_idx=True)
    for qw in (1, 2):
        for ncw in (8, 12):
            L.pdae_tune_patchify(2, qw, ncw)
            got = ops.fps_group(c, G, M, want_idx=True)
            same = all(torch.equal(a, b) for a, b in zip(got, want))
            rec["fused qw=%d ncw=%d" % (qw, ncw)] = {"us": timed_us(lambda: ops.fps_group(c, G, M)), "bit_identical": same}
    out["B=%d N=%d G=%d M=%d" % (B, N, G, M)] = rec
    print(B, N, G, M, json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_patchify.json", "w"), indent=1)
