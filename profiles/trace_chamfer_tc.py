"""Pipeline timeline of the tensor-core Chamfer kernel: clock64 stamps of CTA 0's first tiles (pdae_chamfer_tc_probe)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, synth  # noqa: E402

dev = torch.device("cuda:0")
L = _native.lib()
c = synth.clouds(128, 2048, seed=1)
a, b = torch.from_numpy(synth.prediction(c, seed=1)).to(dev), torch.from_numpy(c).to(dev)
d1 = torch.empty((128, 2048), device=dev); d2 = torch.empty_like(d1)
i1 = torch.empty((128, 2048), dtype=torch.int32, device=dev); i2 = torch.empty_like(i1)
for mode in (3, 2):
    L.pdae_tune_chamfer_tc(mode, 2.0 ** -16)
    for rep in range(2):
        st = torch.zeros(4, dtype=torch.int64, device=dev)
        tr = torch.zeros(256 * 6 + 64 * 4 + 64 * 4, dtype=torch.int64, device=dev)
        rc = L.pdae_chamfer_tc_probe(a.data_ptr(), b.data_ptr(), 128, 2048, 2048, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(),
                                     i2.data_ptr(), None, tr.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _native.check(rc, "probe")
        torch.cuda.synchronize()
    tv = tr.cpu()[1536:1792].view(64, 4)
    tw = tr.cpu()[1792:].view(64, 4)
    t = tr.cpu()[:1536].view(256, 6)
    t0 = int(t[0, 0])
    print("mode", mode, "tile: prod_free prod_committed | epi_wait_start epi_ready epi_released epi_done  (cycles from the first)")
    for k in range(0, 24):
        print(k, [int(x) - t0 for x in t[k]])
    for k in range(1, 256):
        gap = int(t[k, 5]) - int(t[k - 1, 5])
        if gap > 1200:
            print("gap before tile", k, gap, [int(x) - t0 for x in t[k]])
    print("verifier warp 0, row block: start, lists ready, sub 0 done, sub 1 done; epilogue's last tile of that block done")
    for r in range(0, 20):
        print(r, [int(x) - t0 for x in tv[r]], int(t[min(255, r * 8 + 7), 5]) - t0, 'first sub: filtered, main eval, extra rounds:', [int(x) - int(tv[r][1]) for x in tw[r][:3]])
    per = (int(t[200, 5]) - int(t[40, 5])) / 160.0
    print("cycles per tile (tiles 40..200):", per)

from pointdae_b200 import ops  # noqa: E402
pool = []
for s in range(6):
    cc = synth.clouds(128, 2048, seed=100 + s)
    pool.append((torch.from_numpy(synth.prediction(cc, seed=100 + s)).to(dev), torch.from_numpy(cc).to(dev)))
for mode in (0, 2, 3):
    L.pdae_tune_chamfer_tc(mode, 2.0 ** -16)
    for i in range(3):
        ops.chamfer_forward(*pool[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(24):
        ops.chamfer_forward(*pool[i % 6])
    e1.record()
    torch.cuda.synchronize()
    print("mode", mode, "forward us", e0.elapsed_time(e1) / 24 * 1e3)
