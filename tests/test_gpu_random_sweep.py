"""Seeded random sweep over small and awkward shapes: every op through the C ABI against the CPU oracle.
Complements the fixed cases of test_gpu_parity.py (sizes around warp / tile / vector-width boundaries, k = 1,
k = n, single points, duplicated points everywhere).  Indices bit-exact, Chamfer gradients 1e-5."""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from pointdae_b200 import knn_cuda, ops, pointnet2_utils, synth

DEV = "cuda:0"
RNG = np.random.default_rng(20260117)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _cloud(b, n, seed, dup):
    x = synth.clouds(b, n, seed=seed)
    if dup and n >= 8:
        x = synth.adversarial(x, seed=seed, n_small=min(4, n // 8), n_dup=n // 4)
    return x


def _sizes(count, lo, hi):
    special = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1025]
    out = [s for s in special if lo <= s <= hi]
    out += [int(v) for v in RNG.integers(lo, hi + 1, size=count)]
    return out


FPS_CASES = [(int(RNG.integers(1, 4)), n, max(1, int(RNG.integers(1, n + 1))), bool(i % 2)) for i, n in enumerate(_sizes(12, 1, 3000))]


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,m,dup", FPS_CASES)
def test_fps_random(b, n, m, dup):
    xyz = _cloud(b, n, 1000 + n, dup)
    got = pointnet2_utils.furthest_point_sample(cu(xyz), m).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.fps(xyz, m))


KNN_CASES = []
for i, r in enumerate(_sizes(10, 1, 2500)):
    k = int(RNG.integers(1, min(r, 128) + 1))
    KNN_CASES.append((int(RNG.integers(1, 3)), r, int(RNG.integers(1, 70)), 3 if i % 3 else int(RNG.integers(1, 9)), k, bool(i % 2)))
KNN_CASES += [(2, 64, 5, 3, 64, True), (1, 128, 3, 3, 128, False), (2, 40, 9, 3, 1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("b,r,q,dim,k,dup", KNN_CASES)
def test_knn_random(b, r, q, dim, k, dup):
    ref = _cloud(b, r, 2000 + r, dup)
    if dim != 3:
        ref = RNG.standard_normal((b, r, dim)).astype(np.float32) if dim < 3 else np.concatenate(
            [ref, RNG.standard_normal((b, r, dim - 3)).astype(np.float32)], axis=2)
    query = ref[:, RNG.integers(0, r, size=q)].copy()
    query[:, ::2] += np.float32(0.003)
    wd, wi = oracle.knn(ref, query, k)
    D, I = knn_cuda.KNN(k=k, transpose_mode=True)(cu(ref), cu(query))
    np.testing.assert_array_equal(I.cpu().numpy(), wi)
    np.testing.assert_array_equal(D.cpu().numpy(), wd)


CHAMFER_CASES = [(int(RNG.integers(1, 5)), n, int(RNG.integers(1, 1400)), bool(i % 2)) for i, n in enumerate(_sizes(10, 1, 1400))]


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,m,dup", CHAMFER_CASES)
def test_chamfer_random(b, n, m, dup):
    x1, x2 = _cloud(b, n, 3000 + n, dup), _cloud(b, m, 4000 + m, dup)
    if dup:
        x2[:, : min(n, m) // 2] = x1[:, : min(n, m) // 2]  # exact zeros and ties across the clouds
    wd1, wd2, wi1, wi2 = oracle.chamfer_fwd(x1, x2)
    d1, d2, i1, i2 = ops.chamfer_forward(cu(x1), cu(x2))
    np.testing.assert_array_equal(i1.cpu().numpy(), wi1)
    np.testing.assert_array_equal(i2.cpu().numpy(), wi2)
    np.testing.assert_array_equal(d1.cpu().numpy(), wd1)
    np.testing.assert_array_equal(d2.cpu().numpy(), wd2)
    g1 = RNG.uniform(0.5, 1.5, wd1.shape).astype(np.float32)
    g2 = RNG.uniform(0.5, 1.5, wd2.shape).astype(np.float32)
    wg1, wg2 = oracle.chamfer_bwd(x1, x2, wi1, wi2, g1, g2)
    gx1, gx2 = ops.chamfer_backward(cu(x1), cu(x2), i1, i2, cu(g1), cu(g2))
    for got, want in ((gx1, wg1), (gx2, wg2)):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * max(np.abs(want).max(), 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(1, 1), (2, 3), (33, 2), (300, 1), (129, 257), (1000, 40)])
def test_three_nn_random(n, m):
    u, k = _cloud(2, n, 5000 + n, False), _cloud(2, m, 6000 + m, m >= 8)
    wd, wi = oracle.three_nn(u, k)
    d, i = ops.three_nn(cu(u), cu(k))
    np.testing.assert_array_equal(i.cpu().numpy(), wi)
    np.testing.assert_array_equal(d.cpu().numpy(), wd)
