import os, sys, torch
sys.path.insert(0, "/root/repo")
from pointdae_b200 import _native, ops, synth
dev = torch.device("cuda:0"); L = _native.lib()
B, N, G, M = 128, 2048, 64, 32
c = torch.from_numpy(synth.clouds(B, N, seed=N)).to(dev)
def timed_us(fn, reps=40):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [fn() for _ in range(4)]
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (4 * reps)
for qw, ncw in ((1, 12), (1, 8), (2, 12)):
    L.pdae_tune_patchify(2, qw, ncw)
    print(os.environ.get("PDAE_PATCHIFY_DBG"), qw, ncw, timed_us(lambda: ops.fps_group(c, G, M)))
