"""Does chunking the batch over streams inside one forward help?  The scan is FMA-bound, the column recovery is
issue/latency-bound: with the batch in C chunks on C streams the recovery of chunk i can run under the scan of chunk
i+1, and the scan's wave tail is filled by the next chunk's CTAs.  Times graphs of the loss branch (forward, fused
loss, backward) for C = 1, 2, 4."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointdae_b200 import ops, synth

dev = torch.device("cuda:0")
B, N, POOL = 128, 2048, 24
base = torch.from_numpy(synth.clouds(B, N, seed=1)).to(dev)
gen = torch.Generator(device="cpu").manual_seed(1)
clouds = [base[torch.randperm(B, generator=gen).to(dev)][:, torch.randperm(N, generator=gen).to(dev)].contiguous() for _ in range(POOL)]
preds = [c + 0.02 * torch.randn_like(c) for c in clouds]
gone = torch.ones(1, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]


def fwd_chunked(i, C):
    main = torch.cuda.current_stream()
    outs = []
    step = B // C
    for ci in range(C):
        st = main if ci == 0 else streams[ci]
        if st is not main:
            st.wait_stream(main)
        with torch.cuda.stream(st):
            outs.append(ops.chamfer_forward(preds[i][ci * step:(ci + 1) * step], clouds[i][ci * step:(ci + 1) * step]))
    for ci in range(1, C):
        main.wait_stream(streams[ci])
    return outs


def branch(i, C):
    outs = fwd_chunked(i, C)
    d1 = torch.cat([o[0] for o in outs]) if C > 1 else outs[0][0]
    return outs, d1


def time_graphs(fn):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    gs = []
    for i in range(POOL):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn(i)
        gs.append((g, keep))
    for g, _ in gs: g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for r in range(4):
            for g, _ in gs: g.replay()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / (4 * POOL))
    return round(sorted(ts)[2], 2)

out = {"forward_only_C%d_us" % C: time_graphs(lambda i, C=C: fwd_chunked(i, C)) for C in (1, 2, 4)}
print(json.dumps(out))
