"""Kernel timeline of one replayed step graph (torch.profiler / CUPTI): start offsets, durations and streams, to see
where the step's time goes beyond the sum of kernel durations (gaps, overlap between the two branches)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from pointdae_b200 import graphs, ops, synth

PRIO = len(sys.argv) > 1 and sys.argv[1] == "prio"
DEFER = len(sys.argv) > 1 and sys.argv[1] in ("defer", "fused")  # patchifier branch forked after the forward instead of before
FUSED = len(sys.argv) > 1 and sys.argv[1] == "fused"  # bench.py's default step: single-launch patchifier, loss on a third stream

dev = torch.device("cuda:0")
B, N, G, M, POOL = 128, 2048, 64, 32, 8
base = torch.from_numpy(synth.clouds(B, N, seed=1)).to(dev)
gen = torch.Generator(device="cpu").manual_seed(1)
clouds = [base[torch.randperm(B, generator=gen).to(dev)][:, torch.randperm(N, generator=gen).to(dev)].contiguous() for _ in range(POOL)]
preds = [c + 0.02 * torch.randn_like(c) for c in clouds]
gone = torch.ones(1, device=dev)
side = torch.cuda.Stream()
aux = torch.cuda.Stream()

def step(i):
    main = torch.cuda.current_stream()
    if FUSED:
        d1, d2, i1, i2 = ops.chamfer_forward(preds[i], clouds[i])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            nb = ops.fps_group(clouds[i], G, M)[2]
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            l = ops.chamfer_mean_loss(d1, d2)
        g = ops.chamfer_loss_backward(preds[i], clouds[i], i1, i2, d1, d2, gone, 1.0, 1.0)
        main.wait_stream(side)
        main.wait_stream(aux)
        return nb, l, g
    if not DEFER:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, cen = ops.fps_gather(clouds[i], G)
            nb = ops.group_points_knn(clouds[i], cen, M, want_idx=False)
    d1, d2, i1, i2 = ops.chamfer_forward(preds[i], clouds[i])
    if DEFER:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _, cen = ops.fps_gather(clouds[i], G)
            nb = ops.group_points_knn(clouds[i], cen, M, want_idx=False)
    l = ops.chamfer_mean_loss(d1, d2)
    g = ops.chamfer_loss_backward(preds[i], clouds[i], i1, i2, d1, d2, gone, 1.0, 1.0)
    main.wait_stream(side)
    return nb, l, g

for i in range(3): step(i)
torch.cuda.synchronize()
gs = []
for i in range(POOL):
    g = graphs.PriorityGraph() if PRIO else torch.cuda.CUDAGraph()
    with (g.capture() if PRIO else torch.cuda.graph(g)):
        keep = step(i)
    gs.append((g, keep))
for _ in range(3):
    for g, _ in gs: g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for g, _ in gs: g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
rows = [{"name": e.name[:60], "start_us": round(e.time_range.start - t0, 1), "dur_us": round(e.time_range.end - e.time_range.start, 1)} for e in evs]
print(json.dumps(rows[: 14 * 3], indent=0))
