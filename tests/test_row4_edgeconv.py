"""SURVEY.md 8f row 4, oracle stage: one eval-mode EdgeConv layer (graph feature -> 1x1 conv -> BatchNorm -> LeakyReLU ->
max over k, models/dgcnn_util.py:114-126) against the four layers of the reference's OWN `dgcnn_encoder`
(tests/golden/edgeconv_ref.npz, tests/golden/make_golden_edgeconv.py).  Also pins the algebra the kernel will use:
W [x_j - x_i; x_i] = W1 x_j + (W2 - W1) x_i and a per-channel monotone BatchNorm + LeakyReLU turn conv -> max into two small
GEMMs and a gather-max (gather-min where the folded scale is negative).  Tolerance 1e-5 of the output scale: the
convolution's summation order is the library's."""
import os

import numpy as np
import pytest

from oracle import cpu as oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "edgeconv_ref.npz"))


@pytest.mark.parametrize("layer", range(4))
def test_edge_conv_oracle_matches_the_reference_encoder_layer(layer):
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith("l%d/" % layer)}
    got = oracle.edge_conv_max(g["x"], g["idx"], g["weight"], g["scale"], g["shift"], 0.2)
    scale = float(np.abs(g["out"]).max())
    assert got.shape == g["out"].shape
    assert np.allclose(got, g["out"], rtol=1e-5, atol=1e-5 * scale), float(np.abs(got - g["out"]).max())


@pytest.mark.parametrize("layer", range(4))
def test_two_gemms_plus_gather_extremum_equal_the_layer(layer):
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith("l%d/" % layer)}
    x, idx, w, s, t = g["x"], g["idx"], g["weight"], g["scale"], g["shift"]
    b, c, n = x.shape
    w1, w2 = w[:, :c], w[:, c:]
    p = np.einsum("oc,bcn->bon", w1, x).astype(np.float32)             # (B, Co, N)
    q = np.einsum("oc,bcn->bon", w2 - w1, x).astype(np.float32)
    gathered = np.stack([p[bi][:, idx[bi]] for bi in range(b)])          # (B, Co, N, k)
    ext = np.where((s >= 0)[None, :, None], gathered.max(axis=3), gathered.min(axis=3))
    y = (ext + q) * s[None, :, None] + t[None, :, None]
    y = np.where(y >= 0, y, 0.2 * y).astype(np.float32)
    scale = float(np.abs(g["out"]).max())
    assert np.allclose(y, g["out"], rtol=1e-5, atol=2e-5 * scale), float(np.abs(y - g["out"]).max())


def test_gather_half_oracle_composes_to_the_layer():
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith("l2/")}
    x, w = g["x"], g["weight"]
    c = x.shape[1]
    p = np.einsum("bcn,oc->bno", x, w[:, :c]).astype(np.float32)
    q = np.einsum("bcn,oc->bno", x, w[:, c:] - w[:, :c]).astype(np.float32)
    got = oracle.edge_gather_extremum(p, q, g["idx"], g["scale"], g["shift"], 0.2)
    scale = float(np.abs(g["out"]).max())
    assert np.allclose(got, g["out"], rtol=1e-5, atol=2e-5 * scale)


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,k,co", [(2, 200, 20, 64), (1, 33, 5, 7), (3, 1024, 20, 256), (2, 97, 9, 300), (1, 2048, 20, 128)])
def test_gpu_gather_extremum_kernel_is_bit_exact_with_its_oracle(b, n, k, co):
    import torch
    from pointdae_b200 import ops
    rng = np.random.default_rng(n + co)
    p, q = rng.standard_normal((b, n, co)).astype(np.float32), rng.standard_normal((b, n, co)).astype(np.float32)
    idx = rng.integers(0, n, size=(b, n, k)).astype(np.int64)
    scale, shift = rng.standard_normal(co).astype(np.float32), rng.standard_normal(co).astype(np.float32)
    scale[0] = 0.0
    want = oracle.edge_gather_extremum(p, q, idx, scale, shift, 0.2)
    dev = "cuda:0"
    got = ops.edge_gather_extremum(*(torch.from_numpy(a).to(dev) for a in (p, q, idx, scale, shift)), 0.2)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("layer", range(4))
def test_gpu_edge_conv_layer_matches_the_reference_encoder(layer):
    import torch
    from pointdae_b200 import ops
    g = {k.split("/", 1)[1]: GOLD[k] for k in GOLD.files if k.startswith("l%d/" % layer)}
    dev = "cuda:0"
    got = ops.edge_conv_max(*(torch.from_numpy(g[k]).to(dev) for k in ("x", "idx", "weight", "scale", "shift")), 0.2)
    scale = float(np.abs(g["out"]).max())
    assert np.allclose(got.cpu().numpy(), g["out"], rtol=1e-5, atol=2e-5 * scale)
