"""clock64 trace of CTA 0 of the single-launch patchifier: when each centre is posted by the FPS warps and how long the
phases of every search task take.  Writes gpurun_out/trace_patchify.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
L = _native.lib()
B, N, G, M = 128, 2048, 64, 32
c = torch.from_numpy(synth.clouds(B, N, seed=N)).to(dev)
lines = []
for qw, ncw in ((2, 12), (4, 12), (1, 12), (2, 8)):
    L.pdae_tune_patchify(2, qw, ncw)
    ops.fps_group(c, G, M)
    buf = torch.zeros(1 + G + 8 * G, dtype=torch.int64, device=dev)
    L.pdae_patchify_trace(buf.data_ptr())
    ops.fps_group(c, G, M)
    torch.cuda.synchronize()
    L.pdae_patchify_trace(None)
    t = buf.cpu().numpy()
    t0 = t[0]
    post = t[1:1 + G] - t0
    lines.append("== qw=%d ncw=%d  (cycles since the CTA started; 1 us ~ 1965 cycles)" % (qw, ncw))
    lines.append("centres posted at: " + " ".join(str(int(v)) for v in post))
    lines.append("FPS cycles per iteration (median): %d" % int(sorted(post[1:] - post[:-1])[G // 2]))
    ntask = (G + qw - 1) // qw
    lines.append("task  warp  ready   start   pass1+tau  pass2  rescan  sort+emit  fallback  end   (durations)")
    for k in range(ntask):
        s = t[1 + G + 8 * k:1 + G + 8 * k + 6] - t0
        ready = post[min((k + 1) * qw - 1, G - 1)]
        lines.append("%4d  %4d  %6d  %6d  %6d  %6d  %6d  %6d  %6d  %6d" % (
            k, k % ncw, ready, s[0], s[1] - s[0], s[2] - s[1], s[3] - s[2], s[4] - s[3], s[5] - s[4], s[5]))
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/trace_patchify.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
