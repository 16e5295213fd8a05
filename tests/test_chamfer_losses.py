"""Every loss class of extensions/chamfer_dist/__init__.py replayed through this repo's mirror
(pointdae_b200.chamfer_dist) against tests/golden/chamfer_losses.npz -- values produced by the REFERENCE's
own Python classes (tests/golden/make_golden_losses.py).

* CPU (`not gpu`): the mirror's host arithmetic with `chamfer.forward/backward` swapped for the oracle
  stand-in -- checks the torch code above the kernels, nothing is shipped that way.
* GPU: the real product path, sm_100a kernels underneath.
Tolerance: 1e-5 relative (BASELINE.json north_star) on values, 1e-5 of the largest gradient entry on gradients.
"""
import os

import numpy as np
import pytest
import torch

import _loss_cases
from pointdae_b200 import chamfer_dist, synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "chamfer_losses.npz"))
CASES = _loss_cases.cases(synth)


def _check(name, device):
    cls_name, arrays = CASES[name]
    for k, v in arrays.items():  # the fixture's inputs are the generator's inputs
        assert np.array_equal(GOLD["%s/in/%s" % (name, k)], v)
    results, grads = _loss_cases.run(getattr(chamfer_dist, cls_name)(), arrays, device)
    n_out = len([k for k in GOLD.files if k.startswith(name + "/out/")])
    assert len(results) == n_out
    for i, r in enumerate(results):
        want = GOLD["%s/out/%d" % (name, i)]
        assert r.shape == want.shape and r.dtype == want.dtype
        if name == "withnormal_visual" and i == 3:
            # angles in degrees from acos(): d(acos x)/dx is unbounded near x = 1, so one ulp of the cosine (CPU vs
            # CUDA libm) moves small angles by far more than 1e-5 relative; compare the cosines instead
            r, want = np.cos(np.deg2rad(r.astype(np.float64))), np.cos(np.deg2rad(want.astype(np.float64)))
            np.testing.assert_allclose(r, want, rtol=0, atol=2e-6)
            continue
        np.testing.assert_allclose(r, want, rtol=1e-5, atol=1e-6 * max(1.0, float(np.abs(want).max())))
    want_grads = {k.split("/")[-1]: GOLD[k] for k in GOLD.files if k.startswith(name + "/grad/")}
    assert set(grads) == set(want_grads)
    for k, g in grads.items():
        w = want_grads[k]
        assert g.shape == w.shape
        np.testing.assert_allclose(g, w, rtol=1e-5, atol=1e-5 * float(np.abs(w).max()) + 1e-12)


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_arithmetic_matches_reference_classes(name, monkeypatch):
    import _oracle_chamfer
    monkeypatch.setattr(chamfer_dist, "chamfer", _oracle_chamfer)
    _check(name, torch.device("cpu"))


def test_module_exports_every_reference_name():
    # names defined by extensions/chamfer_dist/__init__.py (classes at :14,29,53,123,170,206,237,274,312,348,379,397;
    # helpers at :95-120,200-204)
    for n in ("ChamferFunction ChamferDistanceL2 ChamferDistanceL2_corase2fine ChamferDistanceL2_withnormal "
              "ChamferDistanceL2_withnormal_visual ChamferDistanceL2_withnormalL1 "
              "ChamferDistanceL2_withnormal_strict_normalindex ChamferDistanceL2_withnormal_normalindex "
              "ChamferDistanceL2_withnormal_onlynormalindex ChamferDistanceL2_withnormal_strict ChamferDistanceL2_split "
              "ChamferDistanceL1 dis_normalized_l2 dis_normalized_l1 dis_normalized_l2_strict dis_l2 "
              "unoriented_included_angle").split():
        assert hasattr(chamfer_dist, n), n
    # loss modules are built at import time / before .cuda() in the reference's models: no CUDA in constructors
    chamfer_dist.ChamferDistanceL2_corase2fine()


def test_product_path_has_no_cpu_fallback():
    a = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError):
        chamfer_dist.ChamferDistanceL2()(a, a)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_losses_match_reference_classes(name):
    _check(name, torch.device("cuda:0"))


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["dis_l2", "dis_normalized_l2", "dis_normalized_l1", "dis_normalized_l2_strict"])
@pytest.mark.parametrize("d", [3, 1, 6])
def test_gpu_fused_matched_pair_loss_equals_the_torch_expression(metric, d):
    """ops.matched_pair_loss (csrc/pairloss.cu) against the reference's expression -- torch.gather + the metric + mean
    (extensions/chamfer_dist/__init__.py:95-120, 143-146) -- value and both gradients, incl. a zero vector (F.normalize's
    eps clamp) and an orthogonal pair (torch.min's tie)."""
    import torch
    from pointdae_b200 import chamfer_dist, ops
    dev = "cuda:0"
    g = torch.Generator().manual_seed(d * 7 + len(metric))
    bs, n, m = 3, 70, 90
    a0, b0 = torch.randn(bs, n, d, generator=g), torch.randn(bs, m, d, generator=g)
    a0[0, 0] = 0.0
    if d >= 2:
        a0[1, 1] = 0.0
        a0[1, 1, 0] = 1.0
    idx1 = torch.randint(0, m, (bs, n), generator=g, dtype=torch.int32)
    idx2 = torch.randint(0, n, (bs, m), generator=g, dtype=torch.int32)
    if d >= 2:
        b0[1, int(idx1[1, 1])] = 0.0
        b0[1, int(idx1[1, 1]), 1] = 2.0  # orthogonal to a[1,1]: |u-w|^2 == |u+w|^2
    fn = getattr(chamfer_dist, metric)
    ar, br = a0.to(dev).requires_grad_(True), b0.to(dev).requires_grad_(True)
    i1, i2 = idx1.to(dev), idx2.to(dev)
    want = (torch.mean(fn(ar, chamfer_dist._nearest_rows(br, i1, ar))) + torch.mean(fn(br, chamfer_dist._nearest_rows(ar, i2, br))))
    (want * 1.7).backward()
    ao, bo = a0.to(dev).requires_grad_(True), b0.to(dev).requires_grad_(True)
    got = ops.matched_pair_loss(ao, bo, i1, i2, metric)
    assert got is not None
    (got * 1.7).backward()
    assert abs(float(got) - float(want)) <= 1e-6 * abs(float(want))
    for x, y in ((ao.grad, ar.grad), (bo.grad, br.grad)):
        # (one-channel "normals" normalise to +-1: their gradient is zero up to rounding noise on both sides)
        assert torch.allclose(x, y, rtol=1e-4, atol=max(1e-6 * float(y.abs().max()), 1e-7)), float((x - y).abs().max())
