"""Kernel-level timing on the headline shapes (CUDA events, L2-rotating inputs).  Tuning aid, not the bench."""
import os, sys, json, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointdae_b200 import ops, synth

dev = torch.device("cuda:0")
def timeit(fn, n_in, reps=20, warm=3):
    """GPU time per call: `reps` calls captured into one CUDA graph (no host gaps), replayed 5 times."""
    for i in range(warm): fn(i % n_in)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn(i % n_in)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    B, N, G, M = 128, 2048, 64, 32
    pool = 24
    base = torch.from_numpy(synth.clouds(B, N, seed=1)).to(dev)
    gen = torch.Generator(device="cpu").manual_seed(1)
    clouds = [base[torch.randperm(B, generator=gen).to(dev)][:, torch.randperm(N, generator=gen).to(dev)].contiguous() for _ in range(pool)]
    preds = [c + 0.02 * torch.randn_like(c) for c in clouds]
    res = {}
    if what in ("all", "fps"):
        for cfg in (None, "128,16", "256,8", "512,4", "256,16", "128,32"):
            if cfg: os.environ["PDAE_FPS_CFG"] = cfg
            else: os.environ.pop("PDAE_FPS_CFG", None)
            res["fps_2048_64[%s]" % cfg] = timeit(lambda i: ops.fps_gather(clouds[i], G), pool)
        os.environ.pop("PDAE_FPS_CFG", None)
        c1024 = [c[:, :1024].contiguous() for c in clouds]
        for cfg in (None, "128,8", "256,4", "512,2", "128,16"):
            if cfg: os.environ["PDAE_FPS_CFG"] = cfg
            else: os.environ.pop("PDAE_FPS_CFG", None)
            res["fps_1024_64[%s]" % cfg] = timeit(lambda i: ops.fps_gather(c1024[i], G), pool)
        os.environ.pop("PDAE_FPS_CFG", None)
        c4096 = [torch.from_numpy(synth.clouds(B, 4096, seed=40 + i)).to(dev) for i in range(4)]
        for cfg in (None, "128,32", "256,16", "512,8"):
            if cfg: os.environ["PDAE_FPS_CFG"] = cfg
            else: os.environ.pop("PDAE_FPS_CFG", None)
            res["fps_4096_128[%s]" % cfg] = timeit(lambda i: ops.fps_gather(c4096[i], 128), 4, reps=8)
        os.environ.pop("PDAE_FPS_CFG", None)
        c300 = torch.from_numpy(synth.clouds(B, 300, seed=50)).to(dev)
        res["fps_300_64[None]"] = timeit(lambda i: ops.fps_gather(c300, 64), 1)
        big = torch.from_numpy(synth.clouds(148, 8192, seed=5)).to(dev)
        for cfg in (None, "512,16", "1024,8", "256,32"):
            if cfg: os.environ["PDAE_FPS_CFG"] = cfg
            else: os.environ.pop("PDAE_FPS_CFG", None)
            res["fps_8192_512[%s]" % cfg] = timeit(lambda i: ops.fps_gather(big, 512), 1, reps=3)
        os.environ.pop("PDAE_FPS_CFG", None)
    if what in ("all", "knn"):
        centers = [ops.fps_gather(c, G)[1] for c in clouds]
        res["group_2048_64_32"] = timeit(lambda i: ops.group_points_knn(clouds[i], centers[i], M, want_idx=False), pool)
    if what in ("all", "chamfer"):
        res["chamfer_fwd_2048"] = timeit(lambda i: ops.chamfer_forward(preds[i], clouds[i]), pool)
        res["chamfer_fwd_2048_twopass"] = timeit(lambda i: ops.chamfer_forward(preds[i], clouds[i], symmetric=False), pool)
        d1, d2, i1, i2 = ops.chamfer_forward(preds[0], clouds[0])
        g = torch.full_like(d1, 1.0 / d1.numel())
        res["chamfer_bwd_2048"] = timeit(lambda i: ops.chamfer_backward(preds[i], clouds[i], i1, i2, g, g), pool)
        t = torch.from_numpy(synth.clouds(5000, 36, seed=3)).to(dev)
        t2 = t[:, :32].contiguous() + 0.01
        res["chamfer_fwd_tiny_5000x36x32"] = timeit(lambda i: ops.chamfer_forward(t, t2), 1)
        big = torch.from_numpy(synth.clouds(16, 8192, seed=4)).to(dev)
        bigp = big + 0.01 * torch.randn_like(big)
        res["chamfer_fwd_16x8192"] = timeit(lambda i: ops.chamfer_forward(bigp, big), 1, reps=5)
    for k, v in res.items():
        print("%-36s median %9.1f us   min %9.1f us" % (k, v[0], v[1]))
    pairs = 2.0 * B * N * N
    if "chamfer_fwd_2048" in res:
        print("chamfer fwd frac of FMA peak: %.3f (median)" % (pairs * 6 / (res["chamfer_fwd_2048"][0] * 1e-6) / (148 * 128 * 1.965e9)))
    if "chamfer_fwd_16x8192" in res:
        print("chamfer 16x8192 frac: %.3f" % (2.0 * 16 * 8192 * 8192 * 6 / (res["chamfer_fwd_16x8192"][0] * 1e-6) / (148 * 128 * 1.965e9)))

main()
