"""Times every CTA-shape variant of the Chamfer forward (pdae_tune_chamfer_variant) on the headline shape and on
16x8192^2, checking each against variant 0 bit for bit.  Tuning aid, not the bench.
usage: python profiles/tune_chamfer.py [ids...]  -> JSON on stdout"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pointdae_b200 import _native, ops, synth  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n_in, reps=20):
    for i in range(3):
        fn(i % n_in)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % n_in)
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    ts.sort()
    return round(ts[len(ts) // 2], 2)


def main():
    ids = [int(a) for a in sys.argv[1:]] or [0, 5, 6, 7, 8, 9, 10, 11, 100]
    B, N, pool = 128, 2048, 24
    base = torch.from_numpy(synth.clouds(B, N, seed=1)).to(dev)
    gen = torch.Generator(device="cpu").manual_seed(1)
    clouds = [base[torch.randperm(B, generator=gen).to(dev)][:, torch.randperm(N, generator=gen).to(dev)].contiguous()
              for _ in range(pool)]
    preds = [c + 0.02 * torch.randn_like(c) for c in clouds]
    big = torch.from_numpy(synth.clouds(16, 8192, seed=4)).to(dev)
    bigp = big + 0.01 * torch.randn_like(big)
    rag_a = torch.from_numpy(synth.clouds(7, 1300, seed=5)).to(dev)
    rag_b = torch.from_numpy(synth.adversarial(synth.clouds(7, 777, seed=6), seed=6)).to(dev)
    L = _native.lib()
    L.pdae_tune_chamfer_variant(0)
    want = [ops.chamfer_forward(preds[0], clouds[0]), ops.chamfer_forward(rag_a, rag_b), ops.chamfer_forward(rag_b, rag_a)]
    out = {}
    for v in ids:
        L.pdae_tune_chamfer_variant(v)
        got = [ops.chamfer_forward(preds[0], clouds[0]), ops.chamfer_forward(rag_a, rag_b), ops.chamfer_forward(rag_b, rag_a)]
        same = all(torch.equal(x, y) for w, g in zip(want, got) for x, y in zip(w, g))
        out[str(v)] = {"bit_exact_vs_0": same,
                       "h_128x2048_us": timeit(lambda i: ops.chamfer_forward(preds[i], clouds[i]), pool),
                       "c4_16x8192_us": timeit(lambda i: ops.chamfer_forward(bigp, big), 1, reps=5)}
        print(v, out[str(v)], file=sys.stderr)
    L.pdae_tune_chamfer_variant(0)
    print(json.dumps(out))


main()
