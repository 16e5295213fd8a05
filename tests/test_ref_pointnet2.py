"""`pointnet2_utils` drop-in against outputs of the reference's OWN `extensions/pointnet2/pointnet2_utils.py`
(tests/golden/pointnet2_ref.npz, made on CPU by tests/golden/make_golden_pointnet2.py with `pointnet2._ext` served by the
oracle).  The same driver (`run_all`) pushes every case through a module: CPU = this repo's classes and autograd glue
with `ops` swapped for the oracle stand-in (host logic), GPU = the real thing over the sm_100a kernels.  Indices and
gathers exact; gradients (order-free atomic sums) and the interpolation to 1e-5."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import _oracle_ext
from pointdae_b200 import pointnet2_utils

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "pointnet2_ref.npz"))
_spec = importlib.util.spec_from_file_location("make_golden_pointnet2", os.path.join(HERE, "golden", "make_golden_pointnet2.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

TOLERANT = ("grad", "three_interpolate/out", "three_nn/dist")


def check(out):
    assert sorted(out) == sorted(GOLD.files)
    for name in GOLD.files:
        got, want = out[name], GOLD[name]
        assert got.shape == want.shape and got.dtype == want.dtype, name
        if any(t in name for t in TOLERANT):
            scale = max(float(np.abs(want).max()), 1e-30)
            assert np.allclose(got, want, rtol=1e-5, atol=1e-5 * scale), (name, float(np.abs(got - want).max()))
        else:
            np.testing.assert_array_equal(got, want, err_msg=name)


def test_host_classes_over_the_oracle_match_the_reference_module(monkeypatch):
    monkeypatch.setattr(pointnet2_utils, "ops", _oracle_ext)
    check(gen.run_all(pointnet2_utils, "cpu"))


@pytest.mark.gpu
def test_gpu_pointnet2_utils_matches_the_reference_module():
    check(gen.run_all(pointnet2_utils, "cuda:0"))
