"""BASELINE.json configurations at (or near) full size on one B200: parity against the oracle on a slice the CPU
finishes in seconds, plus size-independent properties on the whole problem (self-distance zero with identity argmin,
FPS indices distinct, kNN lists ascending with the query first, sharded == unsharded)."""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from pointdae_b200 import chamfer_dist, dgcnn_util, group, ops, pointnet2_utils, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_config2_transformer_pretrain_shapes():
    """C2: B=128, N=1024, 64x32 groups; fine Chamfer on ~5000 clouds of 36 vs 32 points; coarse 64 vs 64."""
    xyz = synth.clouds(128, 1024, seed=2)
    t = cu(xyz)
    nb, center = group.Group(64, 32)(t)
    want_nb, want_c, _, _ = oracle.group(xyz[:6], 64, 32)
    np.testing.assert_array_equal(nb[:6].cpu().numpy(), want_nb)
    np.testing.assert_array_equal(center[:6].cpu().numpy(), want_c)
    assert (nb[:, :, 0] == 0).all()  # neighbour 0 of a centre is the centre itself
    fine_a = cu(synth.clouds(5000, 36, seed=3))
    fine_b = nb.reshape(-1, 32, 3)[:5000].contiguous()
    d1, d2, i1, i2 = ops.chamfer_forward(fine_a, fine_b)
    w = oracle.chamfer_fwd(fine_a[:300].cpu().numpy(), fine_b[:300].cpu().numpy())
    np.testing.assert_array_equal(i1[:300].cpu().numpy(), w[2])
    np.testing.assert_array_equal(d2[:300].cpu().numpy(), w[1])
    loss = chamfer_dist.ChamferDistanceL2()(center.requires_grad_(True), center.detach())
    assert loss.item() == 0.0


def test_config3_dgcnn_layers():
    """C3: get_graph_feature kNN k=20 at N=2048 for C = 3, 64, 128 (16 clouds per GPU) + ChamferL1."""
    for c in (3, 64, 128):
        x = synth.features(4 if c > 3 else 16, c, 2048, seed=30 + c)
        t = cu(x)
        idx = dgcnn_util.knn(t, 20)
        want, _ = oracle.feat_knn(x[:1], 20)
        np.testing.assert_array_equal(idx[:1].cpu().numpy(), want)
        assert (idx[:, :, 0] == torch.arange(2048, device=DEV)).all()  # self is the nearest
        f = dgcnn_util.get_graph_feature(t, k=20)
        assert tuple(f.shape) == (x.shape[0], 2 * c, 2048, 20)
        np.testing.assert_array_equal(f[:1].cpu().numpy(), oracle.graph_feature(x[:1], want))
    a = synth.clouds(16, 1024, seed=33)
    l1 = chamfer_dist.ChamferDistanceL1()(cu(synth.prediction(a, seed=33)), cu(a))
    wd1, wd2, _, _ = oracle.chamfer_fwd(synth.prediction(a, seed=33), a)
    want = (np.sqrt(wd1.astype(np.float64)).mean() + np.sqrt(wd2.astype(np.float64)).mean()) / 2
    assert abs(l1.item() - want) < 1e-5 * want


def test_config4_dense_reconstruction():
    """C4: N=8192 -> FPS 512 centres, kNN 32, ChamferL2 vs the 8192-point target (B reduced to 32 to keep the
    test short; the kernels are per-cloud, B only scales the grid)."""
    b = 32
    xyz = synth.clouds(b, 8192, seed=4)
    t = cu(xyz)
    idx, center = group.fps(t, 512)
    np.testing.assert_array_equal(idx[:2].cpu().numpy(), oracle.fps(xyz[:2], 512))
    srt = idx.sort(dim=1)[0]
    assert (srt[:, 1:] != srt[:, :-1]).all()  # distinct samples
    nb, kidx = ops.group_points_knn(t, center, 32, want_idx=True)
    wd, wi = oracle.knn(xyz[:2], center[:2].cpu().numpy(), 32)
    np.testing.assert_array_equal(kidx[:2].cpu().numpy(), wi)
    assert (kidx[:, :, 0] == idx.long()).all()
    pred = cu(synth.prediction(xyz, seed=4))
    d1, d2, i1, i2 = ops.chamfer_forward(pred, t)
    w = oracle.chamfer_fwd(pred[:1].cpu().numpy(), xyz[:1])
    np.testing.assert_array_equal(i1[:1].cpu().numpy(), w[2])
    np.testing.assert_array_equal(i2[:1].cpu().numpy(), w[3])
    np.testing.assert_array_equal(d1[:1].cpu().numpy(), w[0])
    s = ops.chamfer_forward(t, t)
    ar = torch.arange(8192, device=DEV, dtype=torch.int32).expand(b, -1)
    assert (s[0] == 0).all() and torch.equal(s[2], ar) and torch.equal(s[3], ar)


def test_config5_scene_scale_single_gpu():
    """C5: N=100 000 -> FPS 2048, kNN 64, Chamfer with the reference set in 8 slices (the per-rank work of the
    8-GPU run, executed back to back on one GPU, min-combined like the all-reduce does)."""
    n = 100000
    xyz = synth.clouds(1, n, seed=5)
    t = cu(xyz)
    idx, center = group.fps(t, 2048)
    np.testing.assert_array_equal(idx.cpu().numpy(), oracle.fps(xyz, 2048))
    nb, kidx = ops.group_points_knn(t, center, 64, want_idx=True)
    wd, wi = oracle.knn(xyz, center.cpu().numpy()[:, :256], 64)
    np.testing.assert_array_equal(kidx[:, :256].cpu().numpy(), wi)
    q = cu(synth.prediction(xyz, seed=5)[:, :20000])  # 20k queries against the 100k-point reference set
    keys = None
    for r in range(8):
        lo, hi = r * n // 8, (r + 1) * n // 8
        k = ops.chamfer_min_keys(q, t[:, lo:hi].contiguous(), lo)
        keys = k if keys is None else torch.minimum(keys, k)
    sd, si = ops.chamfer_unpack_keys(keys)
    d1, d2, i1, i2 = ops.chamfer_forward(q, t)
    assert torch.equal(sd, d1) and torch.equal(si, i1)
    w = oracle.chamfer_fwd(q[:, :2000].cpu().numpy(), xyz)
    np.testing.assert_array_equal(i1[:, :2000].cpu().numpy(), w[2])
    np.testing.assert_array_equal(d1[:, :2000].cpu().numpy(), w[0])
