"""Seeded cases for three_nn / three_interpolate shared by the golden generator and the tests."""
import numpy as np


def nn_cases(synth):
    """name -> (unknown (B,n,3), known (B,m,3))"""
    cases = {}
    cases["fp_1000_300"] = (synth.clouds(2, 1000, seed=61), synth.clouds(2, 300, seed=62))
    big = synth.clouds(2, 2500, seed=63)
    cases["tiles_2500_1100"] = (big, np.ascontiguousarray(big[:, 7:2207:2]))  # known is a subset: exact zeros, 2 tiles
    dup = synth.adversarial(synth.clouds(3, 200, seed=64), seed=64, n_small=0, n_dup=48)
    cases["ties_333_200"] = (synth.clouds(3, 333, seed=65), dup)
    cases["self_ties_200"] = (dup, dup[:, ::-1].copy())
    cases["m2_50"] = (synth.clouds(2, 50, seed=66), synth.clouds(2, 2, seed=67))
    cases["m1_17"] = (synth.clouds(1, 17, seed=68), synth.clouds(1, 1, seed=69))
    cases["n1_m5"] = (synth.clouds(4, 1, seed=70), synth.clouds(4, 5, seed=71))
    return cases


def interp_inputs(synth, name, unknown, known, idx, dist2, c):
    """features (B,c,m), the reference's inverse-distance weights (models/pointnetv2_util / pointnet2_modules
    FP module: 1/(dist+1e-8) normalised) and an upstream gradient (B,c,n)"""
    rng = np.random.default_rng(synth.BASE_SEED + 9000 + len(name) + unknown.shape[1])
    b, m, _ = known.shape
    n = unknown.shape[1]
    feats = rng.standard_normal((b, c, m)).astype(np.float32)
    dist = np.sqrt(np.where(np.isfinite(dist2), dist2, np.float32(1e4))).astype(np.float32)
    recip = (np.float32(1.0) / (dist + np.float32(1e-8))).astype(np.float32)
    recip = np.minimum(recip, np.float32(1e6))
    weight = (recip / recip.sum(axis=2, keepdims=True)).astype(np.float32)
    gout = rng.standard_normal((b, c, n)).astype(np.float32)
    return feats, weight, gout
