"""the affine corruptions between the patchifier and the encoder
(datasets/corrupt_util_tensor.py:59-343, :706-728; call site models/PointCAE_transformer.py:1010-1017).

tests/golden/corrupt.npz holds outputs of the REFERENCE's own functions on seeded host RNGs
(tests/golden/make_golden_corrupt.py).  CPU: this repo's host mirror must draw the same matrices from the same
seeds (and leave the generators at the same position), and the oracle applied to them must reproduce the stored
outputs.  GPU: the sm_100a kernels must equal the oracle bit for bit and the golden outputs to fp32 rounding.
Tolerance for rotations / shears: 1e-5 relative (BASELINE north_star) + 1e-6 absolute -- the reference leaves the
order of the three products of a row-times-matrix to its BLAS; diagonal maps are compared exactly."""
import os
import random

import numpy as np
import pytest
import torch

import _corrupt_cases as cases
from oracle import cpu as oracle
from pointdae_b200 import corrupt_util_tensor as cut

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "corrupt.npz"))
MATRIX_OF = {"corrupt_scale_nonorm": cut.scale_nonorm_matrix, "corrupt_tranlate": cut.tranlate_matrix,
             "corrupt_rotate_360": cut.rotate_360_matrix, "corrupt_rotate_z_360": cut.rotate_z_360_matrix,
             "corrupt_reflection": cut.reflection_matrix, "corrupt_shear": cut.shear_matrix}
EXACT = {"corrupt_scale_nonorm", "corrupt_tranlate", "corrupt_reflection"}  # diagonal matrices
RTOL, ATOL = 1e-5, 1e-6


def _close(got, want, exact=False):
    if exact:
        assert np.array_equal(got, want)
    else:
        assert np.allclose(got, want, rtol=RTOL, atol=ATOL), float(np.abs(got - want).max())


@pytest.mark.parametrize("name", sorted(cases.SINGLE))
def test_single_corruption_matrix_and_oracle_match_reference(name):
    fn, level, b, g, m = cases.SINGLE[name]
    nb, c = cases.inputs(name, b, g, m)
    cases.seed_all(name)
    mats = MATRIX_OF[fn](b, level).unsqueeze(1).numpy()
    assert np.array_equal(cases.next_draws(), GOLD["single/%s/rng_after" % name])
    tn, tc = oracle.affine_points(nb, c, mats)
    _close(tn, GOLD["single/%s/points" % name], fn in EXACT)
    _close(tc, GOLD["single/%s/center" % name], fn in EXACT)


@pytest.mark.parametrize("name", sorted(cases.CHAINS))
def test_corrupt_data_chain_matches_reference(name):
    typ, b, g, m = cases.CHAINS[name]
    nb, c = cases.inputs(name, b, g, m)
    cases.seed_all(name)
    mats = cut.corrupt_stack(b, typ)
    assert np.array_equal(cases.next_draws(), GOLD["chain/%s/rng_after" % name])
    if mats is None:
        assert typ == ["clean"]
        assert np.array_equal(nb, GOLD["chain/%s/points" % name])
        return
    assert 1 <= mats.shape[1] <= 3 * typ.count("affine_r3")
    tn, tc = oracle.affine_points(nb, c, mats.numpy())
    _close(tn, GOLD["chain/%s/points" % name])
    _close(tc, GOLD["chain/%s/center" % name])
    tn1, tc1 = oracle.affine_points(nb[:, :2], c[:, :2], mats.numpy())
    _close(tn1, GOLD["chain/%s/list1_points" % name])
    _close(tc1, GOLD["chain/%s/list1_center" % name])


def test_unreachable_corruptions_fail_like_the_reference():
    with pytest.raises(NameError):  # datasets/corrupt_util_tensor.py:722 reads an unbound `level`
        cut.corrupt_stack(2, ["jitter"])
    with pytest.raises(RuntimeError):  # CPU tensors never fall back
        cut.corrupt_data(torch.zeros(2, 4, 5, 3), torch.zeros(2, 4, 3), type=["affine_r3"])


def test_oracle_group_affine_identities():
    rng = np.random.default_rng(5)
    xyz = rng.uniform(-1, 1, size=(2, 200, 3)).astype(np.float32)
    nb0, c0, _, _ = oracle.group(xyz, 8, 16)
    eye = np.broadcast_to(np.eye(3, dtype=np.float32), (2, 2, 3, 3)).copy()
    nb, c, tnb, tc, _ = oracle.group_affine(xyz, 8, 16, eye)
    assert np.array_equal(c, c0) and np.array_equal(tc, c0)
    assert np.array_equal(nb, (nb0 + c0[:, :, None]) - c0[:, :, None])  # the reference's round trip, not x - c
    assert np.array_equal(tnb, nb)
    mats = rng.standard_normal((2, 3, 3, 3)).astype(np.float32)
    nb2, _, tnb2, tc2, _ = oracle.group_affine(xyz, 8, 16, mats)
    tp, tcc = oracle.affine_points(nb0 + c0[:, :, None], c0, mats)
    assert np.array_equal(tc2, tcc) and np.array_equal(tnb2, tp - tcc[:, :, None]) and np.array_equal(nb2, nb)


# ------------------------------------------------------------------------------------------------------- GPU
def _dev():
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("t", [0, 1, 3, 8])
def test_gpu_affine_points_bit_exact_with_oracle(t):
    from pointdae_b200 import ops
    rng = np.random.default_rng(100 + t)
    for b, g, m in [(1, 1, 1), (3, 7, 5), (128, 64, 32)]:
        nb = rng.standard_normal((b, g, m, 3)).astype(np.float32)
        c = rng.standard_normal((b, g, 3)).astype(np.float32)
        mats = rng.standard_normal((b, t, 3, 3)).astype(np.float32)
        want_p, want_c = oracle.affine_points(nb, c, mats)
        got_p, got_c = ops.affine_points(torch.from_numpy(nb).to(_dev()), torch.from_numpy(c).to(_dev()), torch.from_numpy(mats))
        assert got_p.shape == nb.shape and got_c.shape == c.shape
        assert np.array_equal(got_p.cpu().numpy(), want_p) and np.array_equal(got_c.cpu().numpy(), want_c)


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,g,m,t", [(4, 1024, 64, 32, 3), (2, 2048, 64, 32, 1), (3, 777, 20, 64, 2), (2, 300, 5, 80, 2),
                                       (2, 9000, 16, 32, 3), (2, 512, 8, 16, 0)])
def test_gpu_group_affine_bit_exact_with_oracle(b, n, g, m, t):
    from pointdae_b200 import ops, synth
    xyz = synth.clouds(b, n, seed=b * 1000 + n)
    mats = np.random.default_rng(n).standard_normal((b, t, 3, 3)).astype(np.float32)
    want_nb, want_c, want_tnb, want_tc, want_idx = oracle.group_affine(xyz, g, m, mats)
    x = torch.from_numpy(xyz).to(_dev())
    _, center = ops.fps_gather(x, g)
    nb, tnb, tc, idx = ops.group_affine(x, center, m, torch.from_numpy(mats), want_idx=True)
    assert np.array_equal(center.cpu().numpy(), want_c)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(nb.cpu().numpy(), want_nb)
    assert np.array_equal(tc.cpu().numpy(), want_tc)
    assert np.array_equal(tnb.cpu().numpy(), want_tnb)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.CHAINS))
def test_gpu_corrupt_data_matches_reference_golden(name):
    typ, b, g, m = cases.CHAINS[name]
    nb, c = cases.inputs(name, b, g, m)
    dn, dc = torch.from_numpy(nb).to(_dev()), torch.from_numpy(c).to(_dev())
    cases.seed_all(name)
    tn, tc = cut.corrupt_data(dn, dc, type=typ)
    _close(tn.cpu().numpy(), GOLD["chain/%s/points" % name])
    _close(tc.cpu().numpy(), GOLD["chain/%s/center" % name])
    cases.seed_all(name)
    tl, cl = cut.corrupt_data([dn, dn[:, :2]], [dc, dc[:, :2]], type=typ)  # list form (models/Point_M2AE.py:799)
    _close(tl[1].cpu().numpy(), GOLD["chain/%s/list1_points" % name])
    _close(cl[1].cpu().numpy(), GOLD["chain/%s/list1_center" % name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.SINGLE))
def test_gpu_single_corruptions_match_reference_golden(name):
    fn, level, b, g, m = cases.SINGLE[name]
    nb, c = cases.inputs(name, b, g, m)
    cases.seed_all(name)
    tn, tc = getattr(cut, fn)(torch.from_numpy(nb).to(_dev()), torch.from_numpy(c).to(_dev()), level)
    _close(tn.cpu().numpy(), GOLD["single/%s/points" % name], fn in EXACT)
    _close(tc.cpu().numpy(), GOLD["single/%s/center" % name], fn in EXACT)


@pytest.mark.gpu
def test_gpu_forward_corrupted_equals_the_models_own_sequence():
    """models/PointCAE_transformer.py:1010-1017 spelled out in torch on the same device."""
    from pointdae_b200 import group, synth
    xyz = torch.from_numpy(synth.clouds(16, 2048, seed=3)).to(_dev())
    divider = group.Group(64, 32)
    random.seed(7), np.random.seed(7), torch.manual_seed(7)
    mats = cut.corrupt_stack(16, ["Drop-Patch", "affine_r3"])
    nb, center, tnb, tc = divider.forward_corrupted(xyz, mats=mats)
    neighborhood, center0 = divider(xyz)
    assert torch.equal(center, center0)
    absn = neighborhood + center0.unsqueeze(2)
    tp, tcen = absn, center0
    for s in range(mats.shape[1]):
        R = mats[:, s].to(_dev())
        tp, tcen = torch.matmul(tp, R.unsqueeze(1)), torch.matmul(tcen, R)
    assert torch.equal(nb, absn - center0.unsqueeze(2))
    scale = float(tp.abs().max())
    assert torch.allclose(tc, tcen, rtol=RTOL, atol=ATOL * scale)
    assert torch.allclose(tnb, tp - tcen.unsqueeze(2), rtol=RTOL, atol=4 * ATOL * scale)
    # seeds drive the same draw when no matrices are given
    random.seed(7), np.random.seed(7), torch.manual_seed(7)
    again = divider.forward_corrupted(xyz, corrupt_type=["Drop-Patch", "affine_r3"])
    assert all(torch.equal(a, b) for a, b in zip(again, (nb, center, tnb, tc)))
    clean = divider.forward_corrupted(xyz, corrupt_type=["clean"])
    assert torch.equal(clean[2], clean[0]) and torch.equal(clean[3], clean[1])


@pytest.mark.gpu
def test_gpu_affine_points_gradient():
    from pointdae_b200 import ops
    g = torch.Generator().manual_seed(11)
    nb = torch.randn(5, 6, 7, 3, generator=g).to(_dev()).requires_grad_(True)
    c = torch.randn(5, 6, 3, generator=g).to(_dev()).requires_grad_(True)
    mats = torch.randn(5, 3, 3, 3, generator=g)
    wp, wc = torch.randn(5, 6, 7, 3, generator=g).to(_dev()), torch.randn(5, 6, 3, generator=g).to(_dev())
    tp, tc = ops.affine_points(nb, c, mats)
    ((tp * wp).sum() + (tc * wc).sum()).backward()
    got = nb.grad.clone(), c.grad.clone()
    nb.grad = c.grad = None
    rp, rc = nb, c
    for s in range(3):
        R = mats[:, s].to(_dev())
        rp, rc = torch.matmul(rp, R.unsqueeze(1)), torch.matmul(rc, R)
    ((rp * wp).sum() + (rc * wc).sum()).backward()
    assert torch.allclose(got[0], nb.grad, rtol=1e-4, atol=1e-5) and torch.allclose(got[1], c.grad, rtol=1e-4, atol=1e-5)


# ---- Drop-Patch (datasets/corrupt_util_tensor.py:592-616): golden = the reference's own function over oracle stand-ins
@pytest.mark.parametrize("name", sorted(cases.DROP_PATCH))
def test_drop_patch_recipe_from_the_oracle_matches_reference(name):
    b, n, level = cases.DROP_PATCH[name]
    pc = cases.drop_patch_input(name, b, n)
    cases.seed_all(name)
    if level is None:
        level = random.random() * 4
    prob = level / 10.0 + 0.5
    idx = oracle.group(pc, cut.NUM_GROUP, cut.GROUP_SIZE)[2]
    patches = pc.reshape(b * n, 3)[(idx + np.arange(b).reshape(-1, 1, 1) * n).reshape(-1)].reshape(b, 64, 32, 3)
    mask = (torch.rand(cut.NUM_GROUP) > prob).numpy()
    if mask.sum() == 0:
        mask[0] = True
    assert np.array_equal(cases.next_draws(), GOLD["drop_patch/%s/rng_after" % name])
    assert np.array_equal(patches[:, mask].reshape(b, -1, 3), GOLD["drop_patch/%s/points" % name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.DROP_PATCH))
def test_gpu_drop_patch_matches_reference_golden(name):
    b, n, level = cases.DROP_PATCH[name]
    pc = torch.from_numpy(cases.drop_patch_input(name, b, n)).to(_dev())
    cases.seed_all(name)
    kept = cut.dropout_patch_random(pc, level)
    assert np.array_equal(cases.next_draws(), GOLD["drop_patch/%s/rng_after" % name])
    assert np.array_equal(kept.cpu().numpy(), GOLD["drop_patch/%s/points" % name])
