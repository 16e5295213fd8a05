"""Deterministic synthetic ShapeNet-shaped clouds shared by tests/, bench.py and the golden-vector
script: each cloud is a mixture -- 50 % on the faces of a random box, 30 % on a
random ellipsoid shell, 20 % in a Gaussian blob -- then centred and scaled into the unit ball like
the reference's pc_norm (datasets/ShapeNet55Dataset.py:67-73).  numpy only (no CUDA, no torch)."""
import numpy as np

BASE_SEED = 20260117


def clouds(b, n, seed=0, dtype=np.float32):
    """-> (b, n, 3) float32, unit-ball normalised."""
    rng = np.random.default_rng(BASE_SEED + int(seed))
    out = np.empty((b, n, 3), dtype=np.float64)
    n_box = n // 2
    n_ell = (3 * n) // 10
    n_blob = n - n_box - n_ell
    for i in range(b):
        half = rng.uniform(0.3, 1.0, size=3)
        pts = rng.uniform(-1.0, 1.0, size=(n_box, 3)) * half
        face = rng.integers(0, 3, size=n_box)
        sign = rng.integers(0, 2, size=n_box) * 2 - 1
        pts[np.arange(n_box), face] = sign * half[face]
        axes = rng.uniform(0.2, 0.9, size=3)
        v = rng.standard_normal(size=(n_ell, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
        ell = v * axes + rng.uniform(-0.3, 0.3, size=3)
        blob = rng.standard_normal(size=(n_blob, 3)) * 0.15 + rng.uniform(-0.5, 0.5, size=3)
        pc = np.concatenate([pts, ell, blob], axis=0)
        pc = pc[rng.permutation(n)]
        pc = pc - pc.mean(axis=0)
        scale = np.max(np.sqrt((pc ** 2).sum(axis=1)))
        pc = pc / scale if scale > 0 else pc + rng.uniform(-0.5, 0.5, size=3)
        out[i] = pc
    return out.astype(dtype)


def prediction(cloud, seed=0, sigma=0.02):
    """Chamfer 'prediction': the cloud with its points permuted plus N(0, sigma^2) noise."""
    rng = np.random.default_rng(BASE_SEED + 7919 + int(seed))
    b, n, _ = cloud.shape
    out = np.empty_like(cloud)
    for i in range(b):
        out[i] = cloud[i][rng.permutation(n)]
    out = out + rng.standard_normal(size=out.shape).astype(np.float32) * np.float32(sigma)
    return out.astype(np.float32)


def adversarial(cloud, seed=0, n_small=8, n_dup=16):
    """Parity-only extras: points with |p|^2 <= 1e-3 incl. exact zeros (FPS skip rule) and exact
    duplicates (FPS / kNN / Chamfer tie rules).  Returns a modified copy."""
    rng = np.random.default_rng(BASE_SEED + 104729 + int(seed))
    out = cloud.copy()
    b, n, _ = out.shape
    for i in range(b):
        pos = rng.permutation(n)
        small = pos[:n_small]
        out[i, small] = (rng.uniform(-0.018, 0.018, size=(n_small, 3))).astype(np.float32)
        if n_small:
            out[i, small[0]] = 0.0
        src = pos[n_small:n_small + n_dup]
        dst = pos[n_small + n_dup:n_small + 2 * n_dup]
        out[i, dst] = out[i, src]
    return out


def features(b, c, n, seed=0):
    """DGCNN-style feature maps (b, c, n): smooth-ish random features so neighbourhoods are meaningful."""
    rng = np.random.default_rng(BASE_SEED + 15485863 + int(seed))
    base = clouds(b, n, seed=seed + 1)  # (b,n,3)
    w = rng.standard_normal(size=(3, c)).astype(np.float32)
    f = np.tanh(base @ w) + 0.05 * rng.standard_normal(size=(b, n, c)).astype(np.float32)
    return np.ascontiguousarray(f.transpose(0, 2, 1)).astype(np.float32)
