"""A/B of the 3-D kNN / Group kernels on the BASELINE shapes (run on the B200 box):

    python profiles/tune_knn.py [--out gpurun_out/r02/tune_knn.json] [--quick]

For every shape: the first-generation kernel (impl 3) is the timing baseline and the bit-exact comparator; the
second-generation kernel (impl 4) is run over a grid of (queries per warp, warps per CTA, TMA on/off, chunks along the
reference cloud) through pdae_tune_knn.  Times are CUDA events around replayed CUDA graphs of `reps` launches on the
launching stream.  Prints one JSON object; `fma_frac` = pairs x 6 lane-ops / time / (148 x 128 x sm clock)."""
import argparse
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointdae_b200 import _native, group, ops, synth  # noqa: E402

DEV = "cuda:0"
PEAK = 148 * 128 * 1.965e9  # lane-ops / s


def timed(fn, reps=10, rounds=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best * 1e3  # us


def tune(impl=4, qw=-1, nw=-1, tile=-1, nz=-1, tma=-1, spec=-1):
    _native.lib().pdae_tune_knn(impl, qw, nw, tile, nz, tma, spec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    shapes = [
        # name, b, r, q, k, planar
        ("H group 128x2048 q64 k32", 128, 2048, 64, 32, False),
        ("C2 group 128x1024 q64 k32", 128, 1024, 64, 32, False),
        ("C4 group 256x8192 q512 k32", 256, 8192, 512, 32, False),
        ("C5 group 1x100000 q2048 k64", 1, 100000, 2048, 64, False),
        ("C3 dgcnn layer-1 knn 16x2048 k20 (planar self)", 16, 2048, 2048, 20, True),
    ]
    out = {}
    for name, b, r, q, k, planar in shapes:
        xyz = torch.from_numpy(synth.clouds(b, r, seed=7 + r)).to(DEV)
        if planar:
            x = xyz.transpose(1, 2).contiguous()
            run = lambda: ops.feat_knn(x, k)
        else:
            _, center = group.fps(xyz, q)
            center = center.contiguous()
            run = lambda: ops.group_points_knn(xyz, center, k, want_idx=True)
        pairs = b * q * r
        rec = {"pairs": pairs, "fma_floor_us": pairs * 6 / PEAK * 1e6, "runs": []}
        tune(impl=3)
        want = run()
        torch.cuda.synchronize()
        t3 = timed(run)
        rec["impl3_us"] = t3
        rec["impl3_fma_frac"] = pairs * 6 / (t3 * 1e-6) / PEAK
        ntiles = (r + 2047) // 2048
        grid = [dict(qw=-1, nw=-1, tma=-1, nz=-1, spec=-1, tile=-1)]
        if not args.quick:
            nzs = [-1] if ntiles == 1 else sorted({1, 2, 3, 4, 6, 8, 12, 16} & set(range(1, ntiles + 1))) + [-1]
            for qw, nw, tma, nz in itertools.product((1, 2, 4), (4, 8), (0, 1), nzs):
                if b * q > 20000 and (qw == 1 or (nz not in (-1, 1))):
                    continue
                grid.append(dict(qw=qw, nw=nw, tma=tma, nz=nz, spec=0, tile=2048))
            if ntiles > 1:  # warp-specialised CTAs (7 compute warps + producer)
                for qw, tma, nz, tile in itertools.product((2, 4), (0, 1), nzs, (1024, 2048)):
                    if b * q > 20000 and nz not in (-1, 1):
                        continue
                    grid.append(dict(qw=qw, nw=8, tma=tma, nz=nz, spec=1, tile=tile))
        for cfg in grid:
            tune(impl=4, **cfg)
            got = run()
            torch.cuda.synchronize()
            same = all((a is None and bb is None) or torch.equal(a, bb) for a, bb in zip(
                got if isinstance(got, tuple) else (got,), want if isinstance(want, tuple) else (want,)))
            t = timed(run)
            rec["runs"].append(dict(cfg, us=t, fma_frac=pairs * 6 / (t * 1e-6) / PEAK, same_as_impl3=bool(same)))
        tune()
        best = min(rec["runs"], key=lambda x: x["us"])
        rec["best"] = best
        rec["auto"] = rec["runs"][0]
        print(name, "impl3 %.1f us | auto %.1f us (%s) | best %.1f us %s" % (
            t3, rec["auto"]["us"], "same" if rec["auto"]["same_as_impl3"] else "DIFFERENT", best["us"],
            {kk: best[kk] for kk in ("qw", "nw", "tma", "nz", "spec", "tile")}), flush=True)
        assert all(x["same_as_impl3"] for x in rec["runs"]), [x for x in rec["runs"] if not x["same_as_impl3"]]
        out[name] = rec
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps({k: {"impl3_us": v["impl3_us"], "auto_us": v["auto"]["us"], "best": v["best"]} for k, v in out.items()}))


if __name__ == "__main__":
    main()
