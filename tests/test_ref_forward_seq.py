"""The first seven lines of the reference model's forward (models/PointCAE_transformer.py:1010-1017) against
tests/golden/forward_seq.npz, which was produced by the reference's own `Group` class statement, its own
`corrupt_data` and the model's own arithmetic (tests/golden/make_golden_forward_seq.py).  CPU: host mirror
(`corrupt_stack`, same seeds) + oracle `group_affine`.  GPU: `Group.forward_corrupted`, two launches.
Clean patches and centres exact; corrupted ones to 1e-5 relative + 1e-6 of the cloud scale (matmul order)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import _corrupt_cases as ccases
from oracle import cpu as oracle
from pointdae_b200 import corrupt_util_tensor as cut

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "forward_seq.npz"))
CASES = {"affine_a": (3, 700, 16, 8, ["affine_r3"]), "affine_b": (2, 1024, 64, 32, ["Drop-Patch", "affine_r3"]),
         "affine_c": (4, 300, 8, 16, ["affine_r3"]), "clean": (2, 256, 8, 8, ["clean"])}


def cloud(name, b, n):
    from pointdae_b200 import synth
    return synth.adversarial(synth.clouds(b, n, seed=sum(map(ord, name))), seed=3)


def check(name, nb, center, tnb, tc):
    np.testing.assert_array_equal(nb, GOLD[name + "/neighborhood"])
    np.testing.assert_array_equal(center, GOLD[name + "/center"])
    scale = float(np.abs(GOLD[name + "/t_center"]).max())
    assert np.allclose(tc, GOLD[name + "/t_center"], rtol=1e-5, atol=1e-6 * scale)
    assert np.allclose(tnb, GOLD[name + "/t_neighborhood"], rtol=1e-5, atol=4e-6 * scale)


def test_generator_and_test_agree_on_the_cases():
    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_golden_forward_seq.py"))
    src = open(spec.origin).read()
    for name, case in CASES.items():
        assert repr(case[-1]) in src.replace('"', "'") and ('"%s"' % name) in src


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_mirror_and_oracle_reproduce_the_models_sequence(name):
    b, n, g, m, typ = CASES[name]
    ccases.seed_all(name)
    mats = cut.corrupt_stack(b, typ)
    assert np.array_equal(ccases.next_draws(), GOLD[name + "/rng_after"])
    mats = np.zeros((b, 0, 3, 3), np.float32) if mats is None else mats.numpy()
    nb, center, tnb, tc, _ = oracle.group_affine(cloud(name, b, n), g, m, mats)
    check(name, nb, center, tnb, tc)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_forward_corrupted_reproduces_the_models_sequence(name):
    from pointdae_b200 import group
    b, n, g, m, typ = CASES[name]
    pts = torch.from_numpy(cloud(name, b, n)).to("cuda:0")
    ccases.seed_all(name)
    nb, center, tnb, tc = group.Group(g, m).forward_corrupted(pts, corrupt_type=typ)
    assert np.array_equal(ccases.next_draws(), GOLD[name + "/rng_after"])
    check(name, nb.cpu().numpy(), center.cpu().numpy(), tnb.cpu().numpy(), tc.cpu().numpy())
