"""Tensor-core 1x1 convolution (csrc/conv_tc.cu: tcgen05.mma kind::tf32, 3xTF32 split, accumulator in tensor memory)
against an fp64 product of the same operands: the split must hold fp32 accuracy (north_star: 1e-5 relative)."""
import pytest
import torch

from pointdae_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [  # b, c, n, j
    (2, 64, 2048, 128), (1, 3, 1000, 128), (2, 128, 1024, 512), (3, 100, 300, 40), (1, 6, 129, 64), (2, 64, 2048, 256),
    (1, 512, 256, 384), (1, 33, 1, 8),
]


@pytest.mark.parametrize("b,c,n,j", SHAPES)
def test_conv1x1_matches_fp64(b, c, n, j):
    g = torch.Generator(device="cpu").manual_seed(c * 1000 + j)
    x = torch.randn(b, c, n, generator=g).to(DEV)
    w = (torch.randn(j, c, generator=g) / c ** 0.5).to(DEV)
    z = ops.conv1x1(x, w)
    want = torch.einsum("jc,bcn->bjn", w.double(), x.double())
    scale = float(want.abs().max())
    err = float((z.double() - want).abs().max())
    assert tuple(z.shape) == (b, j, n)
    assert err <= 1e-5 * scale, (err, scale)  # measured: 1e-7 .. 5e-6 (K = 512) of the largest output
    # the plain fp32 product the reference's layers compute (TF32 off) is no closer to fp64 than this
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref32 = torch.einsum("jc,bcn->bjn", w, x)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert torch.allclose(z, ref32, rtol=1e-5, atol=1e-5 * scale)


def test_conv1x1_extreme_magnitudes_and_exact_small_integers():
    """hi/lo split: values whose low mantissa bits matter (1 + 2^-20) and small integers (exact in tf32)."""
    x = torch.full((1, 32, 128), 1.0 + 2.0 ** -20, device=DEV)
    w = torch.full((16, 32), 1.0 - 2.0 ** -21, device=DEV)
    z = ops.conv1x1(x, w)
    want = 32 * (1.0 + 2.0 ** -20) * (1.0 - 2.0 ** -21)
    assert float((z.double() - want).abs().max()) <= 4e-6, float((z.double() - want).abs().max())
    xi = torch.randint(-8, 9, (2, 40, 200), device=DEV).float()
    wi = torch.randint(-8, 9, (24, 40), device=DEV).float()
    assert torch.equal(ops.conv1x1(xi, wi), torch.einsum("jc,bcn->bjn", wi.double(), xi.double()).float())
