"""Single-launch patchifier (csrc/patchify.cu, `pdae_fps_group_f32`): FPS warps + kNN consumer warps in one CTA per cloud.
Bit-exact against the CPU oracle (`oracle.group` = the reference's Group.forward: misc.fps + KNN + index + subtract,
models/PointCAE_transformer.py:61-86) and against the two-launch path, for every task width / consumer-warp count,
including adversarial clouds (points inside the FPS skip radius, duplicates), mass ties that overflow the candidate
queue (exact warp-select fallback), clouds whose size is not a multiple of 4 or 256, centre counts that are not a
multiple of the task width, and shapes outside the fused range (which must take the two-launch form)."""
import itertools

import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from pointdae_b200 import _native, group, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture
def tune():
    L = _native.lib()

    def set_(enabled=2, qw=1, ncw=8):
        L.pdae_tune_patchify(enabled, qw, ncw)
    yield set_
    L.pdae_tune_patchify(1, 1, 8)


CONFIGS = [dict(qw=qw, ncw=ncw) for qw, ncw in itertools.product((1, 2), (8, 12))]


def _check(xyz, g, m, cfgs, tune):
    want_nb, want_c, want_idx, want_fps = oracle.group(xyz, g, m)
    X = cu(xyz)
    for cfg in cfgs:
        tune(**cfg)
        fps_idx, center, nb, idx = ops.fps_group(X, g, m, want_idx=True)
        np.testing.assert_array_equal(fps_idx.cpu().numpy(), want_fps, err_msg=str(cfg))
        np.testing.assert_array_equal(center.cpu().numpy(), want_c, err_msg=str(cfg))
        np.testing.assert_array_equal(idx.cpu().numpy(), want_idx, err_msg=str(cfg))
        np.testing.assert_array_equal(nb.cpu().numpy(), want_nb, err_msg=str(cfg))
        _, _, nb2, none = ops.fps_group(X, g, m, want_idx=False)
        assert none is None and torch.equal(nb, nb2)


@pytest.mark.parametrize("b,n,g,m,adv", [
    (4, 1024, 64, 32, False), (3, 2048, 64, 32, True), (2, 1000, 33, 17, True), (2, 513, 50, 5, True),
    (2, 2047, 7, 32, True), (2, 512, 64, 1, False), (1, 1500, 1, 32, True), (2, 1025, 130, 20, True), (1, 2048, 1024, 8, False),
])
def test_fused_patchifier_matches_the_oracle_in_every_config(tune, b, n, g, m, adv):
    xyz = synth.clouds(b, n, seed=900 + n + g)
    if adv:
        xyz = synth.adversarial(xyz, seed=n, n_small=min(8, n // 4), n_dup=min(48, n // 4))
    _check(xyz, g, m, CONFIGS, tune)


def test_mass_ties_take_the_exact_fallback(tune):
    """300 copies of each of a few points: every centre sits in a tie class larger than the 64-key queue, so its query
    is redone by the streaming warp-select; the order among equal distances must still be 'lower index first'."""
    b, n, g, m = 2, 2048, 16, 32
    xyz = synth.clouds(b, n, seed=77)
    rng = np.random.default_rng(3)
    for bi in range(b):
        src = rng.integers(0, n, size=6)
        for s in src:
            xyz[bi, rng.integers(0, n, size=300)] = xyz[bi, s]
    _check(xyz, g, m, CONFIGS, tune)


def test_nan_and_inf_points_behave_like_the_two_launch_path(tune):
    """Non-finite coordinates (the CPU oracle's NaN bit patterns differ from the GPU's, so the arbiter here is the
    two-launch path, itself pinned by the other tests): a NaN point keeps the FPS start value 1e10 and is sampled first;
    its search has no finite threshold and goes through the exact fallback; NaN distances order after +inf."""
    b, n, g, m = 2, 1024, 16, 32
    xyz = synth.clouds(b, n, seed=12)
    xyz[0, 5::97] = np.nan
    xyz[1, 7::89, 1] = np.inf
    X = cu(xyz)
    tune(enabled=0)
    want = ops.fps_group(X, g, m, want_idx=True)
    for cfg in CONFIGS:
        tune(**cfg)
        got = ops.fps_group(X, g, m, want_idx=True)
        for a_, b_ in zip(got, want):
            np.testing.assert_array_equal(a_.cpu().numpy(), b_.cpu().numpy(), err_msg=str(cfg))


@pytest.mark.parametrize("b,n,g,m", [(128, 2048, 64, 32), (128, 1024, 64, 32), (16, 2048, 128, 32)])
def test_fused_equals_two_launch_path_at_full_size(tune, b, n, g, m):
    """BASELINE configs H and C2 at full batch: fused launch == fps_gather + group_points_knn, every output."""
    X = cu(synth.clouds(b, n, seed=n + b))
    tune(enabled=0)
    f0, c0, n0, i0 = ops.fps_group(X, g, m, want_idx=True)
    fa, ca = ops.fps_gather(X, g)
    na, ia = ops.group_points_knn(X, ca, m, want_idx=True)
    assert torch.equal(f0, fa) and torch.equal(c0, ca) and torch.equal(n0, na) and torch.equal(i0, ia)
    for cfg in CONFIGS:
        tune(**cfg)
        f1, c1, n1, i1 = ops.fps_group(X, g, m, want_idx=True)
        assert torch.equal(f0, f1) and torch.equal(c0, c1) and torch.equal(n0, n1) and torch.equal(i0, i1), cfg


@pytest.mark.parametrize("b,n,g,m", [(2, 300, 20, 16), (1, 4096, 64, 32), (2, 1024, 16, 48), (1, 8192, 512, 32)])
def test_shapes_outside_the_fused_range_take_the_two_launch_form(tune, b, n, g, m):
    tune()
    xyz = synth.adversarial(synth.clouds(b, n, seed=n), seed=n)
    want_nb, want_c, want_idx, want_fps = oracle.group(xyz, g, m)
    fps_idx, center, nb, idx = ops.fps_group(cu(xyz), g, m, want_idx=True)
    np.testing.assert_array_equal(fps_idx.cpu().numpy(), want_fps)
    np.testing.assert_array_equal(center.cpu().numpy(), want_c)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(nb.cpu().numpy(), want_nb)


def test_group_modules_use_the_fused_launch_and_keep_their_outputs(tune):
    tune()  # force the single launch for this small batch
    xyz = synth.clouds(4, 1024, seed=31)
    want_nb, want_c, want_idx, _ = oracle.group(xyz, 64, 32)
    nb, c = group.Group(64, 32)(cu(xyz))
    np.testing.assert_array_equal(nb.cpu().numpy(), want_nb)
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)
    nb3, c3, flat = group.GroupWithIndex(64, 32)(cu(xyz))
    np.testing.assert_array_equal(nb3.cpu().numpy(), want_nb)
    np.testing.assert_array_equal(flat.cpu().numpy(), (want_idx + np.arange(4).reshape(-1, 1, 1) * 1024).reshape(-1))
    # a cloud that requires grad keeps the differentiable route (same values)
    xg = cu(xyz).requires_grad_(True)
    nbg, cg = group.Group(64, 32)(xg)
    assert nbg.requires_grad and torch.equal(nbg.detach(), nb) and torch.equal(cg.detach(), c)


@pytest.mark.parametrize("b,n,g,m,t", [(4, 1024, 64, 32, 3), (2, 2048, 64, 32, 1), (3, 777, 20, 17, 2), (2, 512, 8, 16, 0),
                                       (2, 300, 5, 16, 2), (1, 4096, 16, 32, 3)])
def test_fused_affine_patchifier_matches_the_oracle(tune, b, n, g, m, t):
    """`pdae_fps_group_affine_f32` = the first seven lines of the reference model's forward
    (models/PointCAE_transformer.py:1010-1017): clean patches, centres, corrupted patches and centres, bit for bit, in the
    single-launch form (forced on these small batches) and, outside its range, in the two-launch form."""
    xyz = synth.adversarial(synth.clouds(b, n, seed=b * 1000 + n), seed=n)
    mats = np.random.default_rng(n).standard_normal((b, t, 3, 3)).astype(np.float32)
    want_nb, want_c, want_tnb, want_tc, want_idx = oracle.group_affine(xyz, g, m, mats)
    for cfg in CONFIGS + [dict(enabled=0)]:
        tune(**cfg)
        fps_idx, c, nb, tnb, tc, idx = ops.fps_group_affine(cu(xyz), g, m, torch.from_numpy(mats), want_idx=True)
        np.testing.assert_array_equal(fps_idx.cpu().numpy(), oracle.fps(xyz, g), err_msg=str(cfg))
        np.testing.assert_array_equal(c.cpu().numpy(), want_c, err_msg=str(cfg))
        np.testing.assert_array_equal(idx.cpu().numpy(), want_idx, err_msg=str(cfg))
        np.testing.assert_array_equal(nb.cpu().numpy(), want_nb, err_msg=str(cfg))
        np.testing.assert_array_equal(tc.cpu().numpy(), want_tc, err_msg=str(cfg))
        np.testing.assert_array_equal(tnb.cpu().numpy(), want_tnb, err_msg=str(cfg))


def test_forward_corrupted_at_full_batch_equals_the_two_launch_form(tune):
    b, n, g, m = 128, 2048, 64, 32
    X = cu(synth.clouds(b, n, seed=5))
    mats = torch.from_numpy(np.random.default_rng(1).standard_normal((b, 3, 3, 3)).astype(np.float32))
    tune(enabled=0)
    want = group.Group(g, m).forward_corrupted(X, mats=mats)
    tune(enabled=1)
    got = group.Group(g, m).forward_corrupted(X, mats=mats)
    for a_, b_ in zip(got, want):
        assert torch.equal(a_, b_)


_RNG = np.random.default_rng(20261018)
RANDOM_CASES = [(int(_RNG.integers(1, 4)), int(_RNG.integers(512, 2049)), int(_RNG.integers(1, 97)), int(_RNG.integers(1, 33)),
                 bool(i % 2), int(_RNG.integers(0, 4))) for i in range(24)]


@pytest.mark.parametrize("b,n,g,m,dup,t", RANDOM_CASES)
def test_random_shapes_inside_the_fused_range(tune, b, n, g, m, dup, t):
    """seeded sweep over cloud size, centre count, group size and matrix-chain length: plain and corrupting epilogue"""
    xyz = synth.clouds(b, n, seed=7000 + n)
    if dup:
        xyz = synth.adversarial(xyz, seed=n, n_small=min(4, n // 8), n_dup=n // 4)
    mats = _RNG.standard_normal((b, t, 3, 3)).astype(np.float32)
    want_nb, want_c, want_tnb, want_tc, want_idx = oracle.group_affine(xyz, g, m, mats)
    tune()
    fps_idx, c, nb, tnb, tc, idx = ops.fps_group_affine(cu(xyz), g, m, torch.from_numpy(mats), want_idx=True)
    f2, c2, nb2, idx2 = ops.fps_group(cu(xyz), g, m, want_idx=True)
    np.testing.assert_array_equal(fps_idx.cpu().numpy(), oracle.fps(xyz, g))
    np.testing.assert_array_equal(c.cpu().numpy(), want_c)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(nb.cpu().numpy(), want_nb)
    np.testing.assert_array_equal(tc.cpu().numpy(), want_tc)
    np.testing.assert_array_equal(tnb.cpu().numpy(), want_tnb)
    assert torch.equal(f2, fps_idx) and torch.equal(c2, c) and torch.equal(idx2, idx)
    # the plain epilogue subtracts the centre once, the corrupting one re-adds and subtracts it (the reference's sequence)
    np.testing.assert_array_equal(nb2.cpu().numpy(), oracle.group(xyz, g, m)[0])


def test_invalid_arguments_are_rejected():
    X = cu(synth.clouds(1, 600, seed=1))
    with pytest.raises(RuntimeError):
        ops.fps_group(X, 8, 601)
    with pytest.raises(RuntimeError):
        ops.fps_group(X[..., :2].contiguous(), 8, 4)
    L = _native.lib()
    assert L.pdae_fps_group_f32(None, 1, 600, 8, 4, None, None, None, None, None, 0, None) != 0
