"""Tensor-core Chamfer filter (csrc/chamfer_tc.cu) against the FP32-pipe kernels: bit-exactness over shapes / data kinds,
filter statistics (largest observed error of the approximate group minima, exact groups per row, fallback rows), the
eps_rel sweep (how far the bound can shrink before a result changes = the margin of the default) and forward timings.
    python profiles/probe_chamfer_tc.py [out.json]"""
import json
import sys
import os
import ctypes

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import _native, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
L = _native.lib()


def fwd(a, b):
    return ops.chamfer_forward(a, b)


def probe(a, b):
    bs, n, m = a.size(0), a.size(1), b.size(1)
    d1 = torch.empty((bs, n), device=dev)
    d2 = torch.empty((bs, m), device=dev)
    i1 = torch.empty((bs, n), dtype=torch.int32, device=dev)
    i2 = torch.empty((bs, m), dtype=torch.int32, device=dev)
    st = torch.zeros(4, dtype=torch.int64, device=dev)
    rc = L.pdae_chamfer_tc_probe(a.data_ptr(), b.data_ptr(), bs, n, m, d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                 st.data_ptr(), None, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _native.check(rc, "probe")
    s = st.cpu().numpy()
    err = float(np.array([s[0] & 0xffffffff], dtype=np.uint32).view(np.float32)[0])
    return [d1, d2, i1, i2], {"max_rel_err": err, "fallback_rows": int(s[1]), "exact_groups": int(s[2]), "rows": int(s[3]),
                              "groups_per_row": float(s[2]) / max(1, int(s[3]))}


def same(x, y):
    return all(torch.equal(p, q) for p, q in zip(x, y))


def mismatches(x, y):
    return int(sum((p != q).sum().item() for p, q in zip(x, y)))


def data(kind, bs, n, m, seed):
    c = synth.clouds(bs, max(n, m), seed=seed)
    if kind == "bench":
        p = synth.prediction(c, seed=seed)
        return p[:, :n].copy(), c[:, :m].copy()
    if kind == "adversarial":
        c2 = synth.adversarial(c, seed=seed, n_small=8, n_dup=200)
        p = synth.prediction(c2, seed=seed, sigma=0.0)  # exact permuted copy: every minimum is 0 and duplicated
        return p[:, :n].copy(), c2[:, :m].copy()
    if kind == "blob":
        rng = np.random.default_rng(seed)
        return (rng.standard_normal((bs, n, 3)) * 0.03).astype(np.float32), c[:, :m].copy()
    if kind == "offset":  # far from the origin: the centring has to carry the precision
        p = synth.prediction(c, seed=seed)
        return (p[:, :n] * 3 + 40).astype(np.float32), (c[:, :m] * 3 + 40).astype(np.float32)
    if kind == "grid":  # lattice: exact ties everywhere
        rng = np.random.default_rng(seed)
        g = rng.integers(0, 12, size=(bs, max(n, m), 3)).astype(np.float32) * 0.125
        return g[:, :n].copy(), g[:, ::-1][:, :m].copy()
    if kind == "same":
        return c[:, :n].copy(), c[:, :m].copy()
    raise ValueError(kind)


out = {"exactness": [], "eps_sweep": [], "timing": {}}
old = L.pdae_tune_chamfer_tc(-1, 0.0)
for kind in ("bench", "adversarial", "blob", "offset", "grid", "same"):
    for (bs, n, m) in ((8, 2048, 2048), (5, 1024, 1024), (3, 2048, 1000), (3, 777, 2041), (2, 512, 2048), (130, 1536, 640)):
        a_np, b_np = data(kind, bs, n, m, seed=hash((kind, n, m)) % 1000)
        a, b = torch.from_numpy(a_np).to(dev), torch.from_numpy(b_np).to(dev)
        L.pdae_tune_chamfer_tc(0, 0.0)
        want = fwd(a, b)
        row = {"kind": kind, "shape": [bs, n, m]}
        for mode in (1, 2, 3):
            L.pdae_tune_chamfer_tc(mode, -1.0)
            got = fwd(a, b)
            row["mode%d_mismatches" % mode] = mismatches(got, want)
        for mode in (2, 3):
            L.pdae_tune_chamfer_tc(mode, -1.0)
            got, st = probe(a, b)
            row["probe%d_mismatches" % mode] = mismatches(got, want)
            row.update({"%s_mode%d" % (k, mode): v for k, v in st.items()})
        out["exactness"].append(row)
        print(json.dumps(row), flush=True)

# eps sweep on the bench shape (mode 1)
a_np, b_np = data("bench", 32, 2048, 2048, seed=11)
a, b = torch.from_numpy(a_np).to(dev), torch.from_numpy(b_np).to(dev)
L.pdae_tune_chamfer_tc(0, 0.0)
want = fwd(a, b)
for mode in (2, 3):
    for e in range(13, 31):
        L.pdae_tune_chamfer_tc(mode, 2.0 ** -e)
        got, st = probe(a, b)
        row = {"mode": mode, "eps_rel": "2^-%d" % e, "mismatches": mismatches(got, want), "groups_per_row": st["groups_per_row"],
               "fallback_rows": st["fallback_rows"], "max_rel_err_log2": float(np.log2(max(st["max_rel_err"], 1e-30)))}
        out["eps_sweep"].append(row)
        print(json.dumps(row), flush=True)
L.pdae_tune_chamfer_tc(old, -1.0)


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (bs, n, m) in ((128, 2048, 2048), (128, 1024, 1024)):
    pool = []
    for s in range(6):  # rotate > L2
        a_np, b_np = data("bench", bs, n, m, seed=100 + s)
        pool.append((torch.from_numpy(a_np).to(dev), torch.from_numpy(b_np).to(dev)))
    it = [0]

    def step():
        a, b = pool[it[0] % len(pool)]
        it[0] += 1
        fwd(a, b)

    row = {}
    for mode in (0, 1, 2, 3):
        L.pdae_tune_chamfer_tc(mode, -1.0)
        row["mode%d_us" % mode] = timeit(step)
    out["timing"]["%dx%dx%d" % (bs, n, m)] = row
    print(json.dumps(row), flush=True)
L.pdae_tune_chamfer_tc(old, -1.0)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
