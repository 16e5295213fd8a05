"""The four `Group` patchifiers of the reference (models/PointCAE_transformer.py:54-86, Point_M2AE_modules.py,
MaskSurf.py, MaskSurf_v2.py).  tests/golden/group_flavours.npz holds what the reference's OWN class statements
return (executed unmodified over oracle-backed stand-ins for KNN / misc.fps / gather_operation,
tests/golden/make_golden_group.py).  CPU: the oracle's `group` restatement composes to the same tensors.
GPU: this repo's fused classes return them bit for bit."""
import os

import numpy as np
import pytest
import torch

import _group_cases as cases
from oracle import cpu as oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "group_flavours.npz"))
PARAMS = [(f, c) for f in sorted(cases.FLAVOURS) for c in sorted(cases.SHAPES)]


def expected_from_oracle(flavour, x, g, m):
    b, n, _ = x.shape
    nb, center, idx, fps_idx = oracle.group(np.ascontiguousarray(x[:, :, :3]), g, m)
    flat = (idx + np.arange(b).reshape(-1, 1, 1) * n).reshape(-1)
    if flavour == "plain":
        return [nb, center]
    if flavour == "with_index":
        return [nb, center, flat]
    if flavour == "normal":
        return [nb, x[:, :, 3:6].reshape(b * n, 3)[flat].reshape(b, g, m, 3), center]
    a = x.shape[2] - 3
    attr = x[:, :, 3:].reshape(b * n, a)
    fflat = (fps_idx.astype(np.int64) + np.arange(b).reshape(-1, 1) * n).reshape(-1)
    return [nb, attr[flat].reshape(b, g, m, a), center, attr[fflat].reshape(b, g, a)]


@pytest.mark.parametrize("flavour,case", PARAMS)
def test_oracle_group_composes_to_the_reference_classes(flavour, case):
    b, n, g, m = cases.SHAPES[case]
    x = cases.inputs(case, b, n, cases.FLAVOURS[flavour][1])
    for i, want in enumerate(expected_from_oracle(flavour, x, g, m)):
        np.testing.assert_array_equal(want, GOLD["%s/%s/%d" % (flavour, case, i)])


@pytest.mark.gpu
@pytest.mark.parametrize("flavour,case", PARAMS)
def test_gpu_group_classes_return_the_reference_tensors(flavour, case):
    from pointdae_b200 import group
    cls = {"plain": group.Group, "with_index": group.GroupWithIndex, "normal": group.GroupNormal,
           "attribute": group.GroupAttribute}[flavour]
    b, n, g, m = cases.SHAPES[case]
    x = cases.inputs(case, b, n, cases.FLAVOURS[flavour][1])
    got = cls(g, m)(torch.from_numpy(x).to("cuda:0"))
    assert len(got) == len([k for k in GOLD.files if k.startswith("%s/%s/" % (flavour, case))])
    for i, t in enumerate(got):
        want = GOLD["%s/%s/%d" % (flavour, case, i)]
        assert t.dtype == torch.from_numpy(want).dtype
        np.testing.assert_array_equal(t.cpu().numpy(), want)
