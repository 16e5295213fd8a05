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

# the same cases through the same reference module over its REAL compiled _ext on a B200 (make_golden_pointnet2.py --gpu):
# what the reference produces on a GPU, incl. torch-CUDA's multiply-by-reciprocal for `/= radius`
GOLD_GPU = np.load(os.path.join(HERE, "golden", "pointnet2_ref_gpu.npz"))

TOLERANT = ("grad", "three_interpolate/out", "three_nn/dist")


def check(out, gold=GOLD):
    assert sorted(out) == sorted(gold.files)
    for name in gold.files:
        got, want = out[name], gold[name]
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
    check(gen.run_all(pointnet2_utils, "cuda:0"), GOLD_GPU)


def test_gpu_made_and_cpu_made_reference_vectors_agree_up_to_the_scalar_division():
    """The two golden files come from the same reference module; they may differ only where torch's CPU and CUDA
    elementwise kernels round differently (division by the radius: 1 ulp) or where atomics reorder a sum."""
    assert sorted(GOLD.files) == sorted(GOLD_GPU.files)
    for name in GOLD.files:
        a, b = GOLD[name], GOLD_GPU[name]
        assert a.shape == b.shape and a.dtype == b.dtype, name
        if a.dtype.kind in "iu":
            np.testing.assert_array_equal(a, b, err_msg=name)
        elif any(t in name for t in TOLERANT) or name.startswith("qag/normalized_ret_xyz"):
            scale = max(float(np.abs(a).max()), 1e-30)
            assert np.allclose(a, b, rtol=1e-5, atol=1e-6 * scale), name
        else:
            np.testing.assert_array_equal(a, b, err_msg=name)


# ---- pointnet2_ops.pointnet2_modules (set abstraction / feature propagation; models/pointnetv2_util.py:317-325) --------
def _modules_forward(dev):
    """A small PointNet++ stack with seeded weights: SA (ball query) -> MSG SA -> global SA -> FP back to the cloud."""
    from pointdae_b200 import pointnet2_modules as p2m
    xyz_np, _, feat_np = gen.cases.inputs()
    torch.manual_seed(5)
    sa1 = p2m.PointnetSAModule(npoint=32, radius=0.3, nsample=8, mlp=[5, 16, 32], use_xyz=True)
    msg = p2m.PointnetSAModuleMSG(npoint=8, radii=[0.4, 0.8], nsamples=[4, 8], mlps=[[32, 16], [32, 24]], use_xyz=True)
    glob = p2m.PointnetSAModule(mlp=[40, 64], use_xyz=True)
    fp = p2m.PointnetFPModule(mlp=[32 + 5, 12])
    mods = [m.to(dev).eval() for m in (sa1, msg, glob, fp)]  # eval: BatchNorm with its initial running statistics
    xyz = torch.from_numpy(xyz_np).to(dev)
    feats = torch.from_numpy(feat_np).to(dev).requires_grad_(True)
    xyz1, f1 = mods[0](xyz, feats)
    xyz2, f2 = mods[1](xyz1, f1)
    none_xyz, f3 = mods[2](xyz2, f2)
    back = mods[3](xyz, xyz1, feats, f1)
    (back.sum() + f3.sum()).backward()
    assert none_xyz is None
    return {"xyz1": xyz1, "f1": f1, "xyz2": xyz2, "f2": f2, "f3": f3, "back": back, "grad": feats.grad,
            "w_grad": mods[0].mlps[0][0].weight.grad}, mods


def test_sa_and_fp_modules_wire_the_ops_like_the_package(monkeypatch):
    from oracle import cpu as oracle
    monkeypatch.setattr(pointnet2_utils, "ops", _oracle_ext)
    out, mods = _modules_forward("cpu")
    xyz_np, _, feat_np = gen.cases.inputs()
    b, n = xyz_np.shape[:2]
    assert sorted(mods[0].state_dict())[:2] == ["mlps.0.0.weight", "mlps.0.1.bias"]
    assert tuple(out["f1"].shape) == (b, 32, 32) and tuple(out["f2"].shape) == (b, 40, 8) and tuple(out["f3"].shape) == (b, 64, 1)
    assert tuple(out["back"].shape) == (b, 12, n) and out["grad"].abs().sum() > 0 and out["w_grad"].abs().sum() > 0
    # first set abstraction spelled out: FPS -> centres, ball query, re-centred xyz in front of the features, MLP, max
    fps_idx = oracle.fps(xyz_np, 32)
    centres = np.take_along_axis(xyz_np, fps_idx[:, :, None].astype(np.int64), axis=1)
    np.testing.assert_array_equal(out["xyz1"].numpy(), centres)
    bq = oracle.ball_query(0.3, 8, xyz_np, centres).astype(np.int64)
    gx = np.stack([xyz_np[i][bq[i]] for i in range(b)]) - centres[:, :, None, :]             # (B, 32, 8, 3)
    gf = np.stack([feat_np[i][:, bq[i]] for i in range(b)])                                   # (B, 5, 32, 8)
    grouped = torch.from_numpy(np.concatenate([gx.transpose(0, 3, 1, 2), gf], axis=1))
    with torch.no_grad():
        want = mods[0].mlps[0](grouped).max(dim=3)[0]
    assert torch.allclose(out["f1"].detach(), want, rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_gpu_sa_and_fp_modules_match_the_cpu_wiring(monkeypatch):
    # the shared MLPs are torch layers: cuDNN runs fp32 convolutions in TF32 by default (1e-4 after two SA levels), which
    # would hide a wiring error of the same size -- compare in true fp32
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    got, _ = _modules_forward("cuda:0")
    monkeypatch.setattr(pointnet2_utils, "ops", _oracle_ext)
    want, _ = _modules_forward("cpu")
    for k in want:
        w, g = want[k].detach(), got[k].detach().cpu()
        scale = float(w.abs().max())
        close = torch.isclose(g, w, rtol=1e-4, atol=1e-5 * scale)
        if k in ("grad", "w_grad"):
            # a max over a group routes its gradient to ONE member: a near-tie may pick another member on another device
            assert (~close).float().mean() < 1e-3, (k, float((~close).float().mean()))
        else:
            assert bool(close.all()), (k, float((g - w).abs().max()))
