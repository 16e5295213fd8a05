"""A/B of the tensor-core forward's CTA shares (PDAE_TCC_BUILD_COST: 0 = equal block counts) on the shapes of the bench:
headline, C2, C4, C5 and one rank's share of the reference-set-sharded C5 forward."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import ops, synth
dev = torch.device("cuda:0")
def t_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {}
for (b, n) in [(128, 2048), (128, 1024), (64, 8192), (1, 100000)]:
    x = torch.from_numpy(synth.clouds(b, n, seed=n)).to(dev); y = x + 0.01 * torch.randn_like(x)
    out["fwd %dx%d" % (b, n)] = t_ms(lambda: ops.chamfer_forward(y, x))
x = torch.from_numpy(synth.clouds(1, 100000, seed=5)).to(dev); y = x + 0.01 * torch.randn_like(x)
for w in (2, 8):
    sl = x[:, : 100000 // w].contiguous()
    out["sharded share world %d" % w] = t_ms(lambda: ops.chamfer_sharded_local(y, sl, 0))
print(os.environ.get("PDAE_TCC_BUILD_COST"), json.dumps(out))
