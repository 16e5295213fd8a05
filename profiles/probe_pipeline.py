"""Software pipeline inside one step: the batch in C chunks, the FMA-bound scan of chunk j+1 ordered behind the scan of
chunk j (event) and running over chunk j's latency-bound tail (column recovery, backward).  Times replayed graphs of the
whole headline step (patchifier on a third stream) for C = 1 (bench.py's step), 2, 4 and several column-split settings."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointdae_b200 import _native, ops, synth

dev = torch.device("cuda:0")
B, N, G, M, POOL = 128, 2048, 64, 32, 24
base = torch.from_numpy(synth.clouds(B, N, seed=1)).to(dev)
gen = torch.Generator(device="cpu").manual_seed(1)
clouds = [base[torch.randperm(B, generator=gen).to(dev)][:, torch.randperm(N, generator=gen).to(dev)].contiguous() for _ in range(POOL)]
preds = [c + 0.02 * torch.randn_like(c) for c in clouds]
gone = torch.ones(1, device=dev)
side = [torch.cuda.Stream() for _ in range(4)]
patch = torch.cuda.Stream()
L = _native.lib()


def step(i, C, patchifier=True, split_in_step=True):
    c, p = clouds[i], preds[i]
    main = torch.cuda.current_stream()
    if patchifier:
        patch.wait_stream(main)
        with torch.cuda.stream(patch):
            _, center = ops.fps_gather(c, G)
            nb, _ = ops.group_points_knn(c, center, M, want_idx=False)
    h = B // C
    outs, grads = [], []
    prev = None
    with ops.chamfer_column_split(split_in_step):
        for j in range(C):
            st = main if j == 0 else side[j]
            if st is not main:
                st.wait_stream(main) if prev is None else None
            with torch.cuda.stream(st):
                if j > 0:
                    st.wait_event(prev)           # scan j starts when scan j-1 is done; its tail runs underneath
                ev = torch.cuda.Event() if j + 1 < C else None
                pj, cj = p[j * h:(j + 1) * h], c[j * h:(j + 1) * h]
                o = ops.chamfer_forward(pj, cj, scan_done=ev)
                g = ops.chamfer_loss_backward(pj, cj, o[2], o[3], o[0], o[1], gone, 1.0 / C, 1.0 / C)
                outs.append(o), grads.append(g)
                prev = ev
    for j in range(1, C):
        main.wait_stream(side[j])
    d1 = torch.cat([o[0] for o in outs]) if C > 1 else outs[0][0]
    d2 = torch.cat([o[1] for o in outs]) if C > 1 else outs[0][1]
    loss = ops.chamfer_mean_loss(d1, d2)[0]
    if patchifier:
        main.wait_stream(patch)
    return loss, grads


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
    return round(sorted(ts)[2], 2), gs


out = {"unit": "us per step, replayed CUDA graphs, median of 5"}
ref_loss = None
for C in (1, 2, 4):
    for nc in (1, 0, 2, 4):
        L.pdae_tune_chamfer_split(nc)
        for patchifier in (True, False):
            t, gs = time_graphs(lambda i, C=C: step(i, C, patchifier))
            loss = float(gs[0][1][0])
            if ref_loss is None:
                ref_loss = loss
            out["C%d_split%s_%s" % (C, "auto" if nc == 0 else nc, "step" if patchifier else "lossbranch")] = [t, abs(loss - ref_loss) <= 1e-6 * abs(ref_loss)]
            del gs
L.pdae_tune_chamfer_split(0)
print(json.dumps(out))
