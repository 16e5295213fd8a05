"""The reference's pure-PyTorch CPU path for the geometry chain, used ONLY as the timed CPU baseline
(bench.py `cpu_baseline` and `--impl reference`) -- test/measurement infrastructure, never imported
by the product package.

Recipe (BASELINE.md 2b / SURVEY.md 8d), each piece following the reference's own pure-torch helper:
  * FPS loop      : segmentation/pointnet_util.py:53-73 (farthest_point_sample), start index fixed to 0
  * Group         : torch.cdist + topk(k, largest=False) + gather + centre-subtract
                    (semantics of models/PointCAE_transformer.py:61-86)
  * Chamfer L2    : broadcast pairwise squared distance (segmentation/pointnet_util.py:22-36 style),
                    min both ways, mean + mean (extensions/chamfer_dist/__init__.py:29-44), autograd backward
It is not a parity oracle (no origin-skip rule, torch tie order); oracle/pdae_oracle.c is.
"""
import time

import torch


def farthest_point_sample(xyz, npoint):
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.ones(B, N) * 1e10
    farthest = torch.zeros(B, dtype=torch.long)
    batch_indices = torch.arange(B, dtype=torch.long)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[batch_indices, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        distance = torch.min(distance, dist)
        farthest = torch.max(distance, -1)[1]
    return centroids


def group(xyz, num_group, group_size):
    B, N, _ = xyz.shape
    fps_idx = farthest_point_sample(xyz, num_group)
    center = torch.gather(xyz, 1, fps_idx[..., None].expand(-1, -1, 3))
    d = torch.cdist(center, xyz)
    idx = d.topk(group_size, dim=-1, largest=False)[1]  # B G M
    idx = idx + torch.arange(B).view(-1, 1, 1) * N
    nb = xyz.reshape(B * N, 3)[idx.view(-1)].view(B, num_group, group_size, 3)
    return nb - center.unsqueeze(2), center


def chamfer_l2(pred, target, chunk=4):
    """pairwise Chamfer L2 with autograd; processed `chunk` clouds at a time to bound memory."""
    tot1 = pred.new_zeros(())
    tot2 = pred.new_zeros(())
    B = pred.size(0)
    for s in range(0, B, chunk):
        p, t = pred[s:s + chunk], target[s:s + chunk]
        d = torch.sum((p[:, :, None] - t[:, None]) ** 2, dim=-1)  # b n m
        tot1 = tot1 + d.min(dim=2)[0].sum()
        tot2 = tot2 + d.min(dim=1)[0].sum()
    return tot1 / (B * pred.size(1)) + tot2 / (B * target.size(1))


def step(cloud, pred, num_group, group_size):
    """One pass of the chain over a batch: FPS -> Group -> Chamfer L2 fwd + bwd.  Returns the loss value."""
    nb, center = group(cloud, num_group, group_size)
    pred = pred.detach().requires_grad_(True)
    loss = chamfer_l2(pred, cloud)
    loss.backward()
    return float(loss.detach()), nb, center, pred.grad


def time_step(cloud, pred, num_group, group_size, reps=3, warmup=1):
    """median seconds of `step` over `reps` repetitions."""
    for _ in range(warmup):
        step(cloud, pred, num_group, group_size)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step(cloud, pred, num_group, group_size)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2]
