"""Small native-step calls for compute-sanitizer memcheck: single-launch patchifier (48 clouds) and the two-launch /
FP32-pipe forms (3 clouds of 700 points), twice each on the same buffers."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import ops, synth
dev = torch.device("cuda:0")
gone = torch.ones(1, device=dev)
for (b, n, g, m) in [(48, 1024, 16, 32), (3, 700, 9, 17)]:
    xyz = synth.adversarial(synth.clouds(b, n, seed=n), seed=n)
    c = torch.from_numpy(xyz).to(dev); p = torch.from_numpy(synth.prediction(xyz, seed=1)).to(dev)
    bufs = ops.StepBuffers(b, n, g, m, dev)
    for _ in range(2):
        o = ops.hot_step(c, p, g, m, gone, buffers=bufs)
    torch.cuda.synchronize()
    f, ce, nb, _ = ops.fps_group(c, g, m)
    assert torch.equal(o.neighborhood, nb) and torch.equal(o.center, ce)
print("sanitize_step ok")
