"""CUDA-event timings of SURVEY.md 8f row 3 at the headline shape (B=128, N=2048, 64x32 patches, chain of 3 matrices):
the model's own sequence (models/PointCAE_transformer.py:1010-1017, torch ops on the GPU, patchifier = this repo's Group
in both arms) against the one-launch chain and the fused Group epilogue.  Prints one JSON object."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import group, ops, synth  # noqa: E402


def timed(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    B, N, G, M, T = 128, 2048, 64, 32, 3
    xyz = torch.from_numpy(synth.clouds(B, N, seed=1)).to(dev)
    g = torch.Generator().manual_seed(0)
    mats = torch.randn(B, T, 3, 3, generator=g)
    dm = mats.to(dev)
    divider = group.Group(G, M)
    nb, center = divider(xyz)

    def torch_sequence(nb, center):
        absn = nb + center.unsqueeze(2)
        tp, tc = absn, center
        for s in range(T):
            R = dm[:, s]
            tp, tc = torch.matmul(tp, R.unsqueeze(1)), torch.matmul(tc, R)
        return absn - center.unsqueeze(2), tp - tc.unsqueeze(2), tc

    def ours_chain(nb, center):
        absn = nb + center.unsqueeze(2)
        tp, tc = ops.affine_points(absn, center, dm)
        return absn - center.unsqueeze(2), tp - tc.unsqueeze(2), tc

    out = {"shape": {"B": B, "N": N, "G": G, "M": M, "T": T}, "unit": "us per call, eager launches, CUDA events"}
    out["group_only"] = timed(lambda: divider(xyz))
    out["group+torch_sequence (reference forward, 10 torch kernels)"] = timed(lambda: torch_sequence(*divider(xyz)))
    out["group+one_launch_chain (drop-in corrupt_data, models unchanged)"] = timed(lambda: ours_chain(*divider(xyz)))
    out["forward_corrupted (fused kNN epilogue)"] = timed(lambda: divider.forward_corrupted(xyz, mats=dm))
    out["torch_sequence_alone"] = timed(lambda: torch_sequence(nb, center))
    absn = nb + center.unsqueeze(2)
    out["affine_points_alone"] = timed(lambda: ops.affine_points(absn, center, dm))
    nbytes = (absn.numel() + center.numel()) * 4 * 2
    out["affine_points_GBps"] = nbytes / (out["affine_points_alone"] * 1e-6) / 1e9
    a, b_, c = torch_sequence(nb, center)
    f = divider.forward_corrupted(xyz, mats=dm)
    out["max_abs_diff_vs_torch"] = [float((f[0] - a).abs().max()), float((f[2] - b_).abs().max()), float((f[3] - c).abs().max())]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
