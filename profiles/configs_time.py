"""Timing of the BASELINE.json configurations at full size on one B200 (CUDA events, 5 reps, median)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pointdae_b200 import ops, synth, dgcnn_util
dev = torch.device("cuda:0")
def t_ms(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
res = {}
def cloud(b, n, seed):
    base = torch.from_numpy(synth.clouds(min(b, 8), n, seed=seed)).to(dev)
    return base.repeat((b + base.size(0) - 1) // base.size(0), 1, 1)[:b].contiguous() + 0.001 * torch.randn(b, n, 3, device=dev)
# C2
c = cloud(128, 1024, 2)
res["C2 fps 1024->64 (B=128)"] = t_ms(lambda: ops.fps_gather(c, 64))
cen = ops.fps_gather(c, 64)[1]
res["C2 group k=32"] = t_ms(lambda: ops.group_points_knn(c, cen, 32, want_idx=False))
fa = cloud(5000, 36, 3); fb = cloud(5000, 32, 4)
res["C2 chamfer fine 5000x36x32 fwd"] = t_ms(lambda: ops.chamfer_forward(fa, fb))
# C3
for C in (3, 64, 128):
    x = torch.from_numpy(synth.features(16, C, 2048, seed=C)).to(dev)
    res["C3 dgcnn knn C=%d N=2048 B=16" % C] = t_ms(lambda: dgcnn_util.knn(x, 20), reps=3)
    idx = dgcnn_util.knn(x, 20)
    res["C3 graph_feature C=%d" % C] = t_ms(lambda: ops._graph_feature_fwd(x, idx), reps=3)
    g = torch.randn(16, 2048, 20, 2 * C, device=dev)
    res["C3 graph_feature backward C=%d" % C] = t_ms(lambda: ops._graph_feature_bwd(g, idx, C, 2048), reps=3)
    del g
# C4
c = cloud(256, 8192, 5)
res["C4 fps 8192->512 (B=256)"] = t_ms(lambda: ops.fps_gather(c, 512), reps=3)
cen = ops.fps_gather(c, 512)[1]
res["C4 group k=32 (Q=512,R=8192)"] = t_ms(lambda: ops.group_points_knn(c, cen, 32, want_idx=False), reps=3)
p = c + 0.01 * torch.randn_like(c)
res["C4 chamfer 256x8192^2 fwd"] = t_ms(lambda: ops.chamfer_forward(p, c), reps=3)
res["C4 chamfer fwd frac of FMA peak (algorithmic)"] = 2.0 * 256 * 8192 * 8192 * 6 / (res["C4 chamfer 256x8192^2 fwd"] * 1e-3) / (148 * 128 * 1.965e9)
d1, d2, i1, i2 = ops.chamfer_forward(p, c); g = torch.full_like(d1, 1e-6)
res["C4 chamfer bwd"] = t_ms(lambda: ops.chamfer_backward(p, c, i1, i2, g, g), reps=3)
del c, p, d1, d2, i1, i2, g
# C5
c = cloud(1, 100000, 6)
res["C5 fps 100k->2048 (B=1)"] = t_ms(lambda: ops.fps_gather(c, 2048), reps=2)
cen = ops.fps_gather(c, 2048)[1]
res["C5 group k=64 (Q=2048,R=100k)"] = t_ms(lambda: ops.group_points_knn(c, cen, 64, want_idx=False), reps=3)
p = c + 0.01 * torch.randn_like(c)
res["C5 chamfer 100k^2 fwd (1 GPU, unsharded)"] = t_ms(lambda: ops.chamfer_forward(p, c), reps=3)
sl = c[:, :12500].contiguous()
res["C5 chamfer 100k x 12.5k slice keys (per-rank work of 8)"] = t_ms(lambda: ops.chamfer_min_keys(p, sl, 0), reps=3)
for k, v in res.items():
    print("%-58s %10.3f %s" % (k, v, "" if "frac" in k else "ms"))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "configs_time.json"), "w"), indent=1)
