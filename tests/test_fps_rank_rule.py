"""CPU check of the closed-form tie rule the CUDA FPS kernels implement against the oracle's literal simulation of
the reference's shared-memory tree (sampling_gpu.cu:62-68,118-171): arg-max of (value, smaller bit-reversed slot,
smaller k) must pick the same point as the tree for every block size, on inputs made almost entirely of ties."""
import numpy as np
import pytest

from oracle import cpu as oracle


def bitrev(v, bits):
    r = 0
    for b in range(bits):
        r |= ((v >> b) & 1) << (bits - 1 - b)
    return r


def rank_rule_fps(xyz, m):
    """numpy restatement of the kernels' rule (fps.cu: fps_rank / fps_val_bits), fp32 with explicit fma order."""
    b, n, _ = xyz.shape
    bs = oracle.fps_block_size(n)
    lg = int(np.log2(bs))
    out = np.zeros((b, m), dtype=np.int32)
    k = np.arange(n)
    rank = np.array([(bitrev(int(kk) & (bs - 1), lg) << 22) | (int(kk) >> lg) for kk in k], dtype=np.int64)
    f = np.float32
    for bi in range(b):
        p = xyz[bi].astype(np.float32)
        mag = (p[:, 2] * p[:, 2]).astype(f)  # exact rounding order is irrelevant for this test's inputs (grid points)
        mag = (p[:, 0].astype(np.float64) ** 2 + p[:, 1].astype(np.float64) ** 2 + p[:, 2].astype(np.float64) ** 2)
        valid = mag > 1e-3
        temp = np.full(n, 1e10, dtype=np.float32)
        old = 0
        for j in range(1, m):
            d = ((p - p[old]).astype(np.float64) ** 2).sum(axis=1).astype(np.float32)
            temp = np.where(valid, np.minimum(d, temp), temp)
            if not valid.any():
                old = 0
            else:
                cand = np.where(valid, temp, -np.inf)
                best = cand.max()
                ties = np.flatnonzero(cand == best)
                old = int(ties[np.argmin(rank[ties])])
            out[bi, j] = old
    return out


@pytest.mark.parametrize("n", [2, 3, 8, 37, 64, 100, 300, 512, 513, 700, 1024, 1500])
def test_rank_rule_equals_reference_tree_on_heavy_ties(n):
    rng = np.random.default_rng(n)
    # points on a coarse integer grid scaled by 0.25: distances are exact in fp32 (no rounding ambiguity) and
    # most candidates tie, so the selection is decided almost purely by the tie rule
    xyz = (rng.integers(-2, 3, size=(3, n, 3)) * 0.25).astype(np.float32)
    m = min(n, 24)
    np.testing.assert_array_equal(rank_rule_fps(xyz, m), oracle.fps(xyz, m))
