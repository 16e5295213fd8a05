"""Generates tests/golden/pointnet2_ref.npz in THIS container (CPU only):

    python tests/golden/make_golden_pointnet2.py

Imports the REFERENCE's own `extensions/pointnet2/pointnet2_utils.py` from /root/reference with `pointnet2._ext` served
by the oracle-backed stand-in tests/_oracle_ext.py (`pytorch_utils` and `ipdb` stubbed: neither is touched on this
path) and stores what its public functions and modules return on seeded inputs: furthest_point_sample,
gather_operation (+grad), three_nn, three_interpolate (+grad), ball_query, grouping_operation (+grad), QueryAndGroup in
five configurations (incl. the seeded `sample_uniformly` resampling), GroupAll.

    python tests/golden/make_golden_pointnet2.py --gpu      (on the B200 box, via gpurun)

writes gpurun_out/pointnet2_ref_gpu.npz (committed as tests/golden/pointnet2_ref_gpu.npz): the same cases through the
same reference module, but over the reference's REAL compiled `_ext` (oracle/_ref/pointnet2_ext/_ext.so, its CUDA rebuilt
unmodified for sm_100a) on cuda:0, the module text read from the copy oracle/build_ref.py stages under oracle/_ref/pysrc/
(git-ignored; /root/reference does not exist on that box).  These are what the reference produces ON A GPU -- torch-CUDA
turns `x /= python_scalar` into a multiply by the reciprocal (pointnet2_utils.py:350-351), which the CPU-made file
cannot show (1 ulp on 2.5 % of the normalised offsets)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle_ext  # noqa: E402
import _pointnet2_cases as cases  # noqa: E402

REF = "/root/reference/extensions/pointnet2/pointnet2_utils.py"


REF_STAGED = os.path.join(ROOT, "oracle", "_ref", "pysrc", "pointnet2_utils.py")
REF_EXT_SO = os.path.join(ROOT, "oracle", "_ref", "pointnet2_ext", "_ext.so")


def load_reference(ext=None, path=None):
    """The reference module, executed over `ext` as its `pointnet2._ext` (default: the oracle-backed stand-in)."""
    ext = _oracle_ext if ext is None else ext
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    sys.modules.setdefault("pytorch_utils", types.ModuleType("pytorch_utils"))
    pkg = types.ModuleType("pointnet2")
    pkg.__path__ = []
    pkg._ext = ext
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = ext
    spec = importlib.util.spec_from_file_location("ref_pointnet2_utils", path or (REF if os.path.exists(REF) else REF_STAGED))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_all(p2u, dev="cpu"):
    """Every case through the module `p2u` (the reference's, or this repo's drop-in): name -> numpy array."""
    xyz_np, new_np, feat_np = cases.inputs()
    xyz, new_xyz = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(new_np).to(dev)
    out = {}

    def keep(name, t):
        out[name] = t.detach().cpu().numpy()

    fps_idx = p2u.furthest_point_sample(xyz, 32)
    keep("fps/idx", fps_idx)
    feats = torch.from_numpy(feat_np).to(dev).requires_grad_(True)
    gathered = p2u.gather_operation(feats, fps_idx)
    keep("gather/out", gathered)
    gathered.backward(cases.weights_for("gather", gathered.shape).to(dev))
    keep("gather/grad", feats.grad)

    dist, idx3 = p2u.three_nn(xyz, new_xyz)
    keep("three_nn/dist", dist)
    keep("three_nn/idx", idx3)
    weight = (1.0 / (dist + 1e-8))
    weight = (weight / weight.sum(dim=2, keepdim=True)).contiguous()
    known = torch.from_numpy(feat_np[:, :, : cases.M].copy()).to(dev).requires_grad_(True)
    interp = p2u.three_interpolate(known, idx3, weight)
    keep("three_interpolate/out", interp)
    interp.backward(cases.weights_for("interp", interp.shape).to(dev))
    keep("three_interpolate/grad", known.grad)

    bq = p2u.ball_query(0.25, 16, xyz, new_xyz)
    keep("ball_query/idx", bq)
    feats2 = torch.from_numpy(feat_np).to(dev).requires_grad_(True)
    grouped = p2u.grouping_operation(feats2, bq)
    keep("grouping/out", grouped)
    grouped.backward(cases.weights_for("grouping", grouped.shape).to(dev))
    keep("grouping/grad", feats2.grad)

    for name, (kw, with_features) in cases.QAG.items():
        torch.manual_seed(1234)  # sample_uniformly draws torch.randint on the default CPU generator
        f = torch.from_numpy(feat_np).to(dev).requires_grad_(True) if with_features else None
        res = cases.as_tuple(p2u.QueryAndGroup(**kw)(xyz, new_xyz, f))
        for i, r in enumerate(res):
            keep("qag/%s/%d" % (name, i), r)
        if f is not None:
            res[0].backward(cases.weights_for("qag" + name, res[0].shape).to(dev))
            keep("qag/%s/grad" % name, f.grad)
    for name, (kw, with_features) in cases.GROUP_ALL.items():
        f = torch.from_numpy(feat_np).to(dev) if with_features else None
        module = p2u.GroupAll(**kw)
        if not hasattr(module, "ret_grouped_xyz"):
            # the reference's forward reads an attribute its constructor never sets (:381-384 vs :421): as shipped it
            # raises AttributeError; set to the constructor's default to obtain the output the code intends
            module.ret_grouped_xyz = False
        res = cases.as_tuple(module(xyz, new_xyz, f))
        keep("group_all/%s/0" % name, res[0])
    return out


def load_real_ext():
    """the reference's compiled extension, rebuilt unmodified by oracle/build_ref.py"""
    spec = importlib.util.spec_from_file_location("_ext", REF_EXT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main_gpu():
    assert torch.cuda.is_available(), "--gpu needs the B200 box"
    out = run_all(load_reference(load_real_ext()), "cuda:0")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "pointnet2_ref_gpu.npz")
    np.savez_compressed(path, **out)
    cpu = np.load(os.path.join(ROOT, "tests", "golden", "pointnet2_ref.npz"))
    for k in sorted(out):
        same = cpu[k].shape == out[k].shape and np.array_equal(cpu[k], out[k])
        print("%-40s %s" % (k, "== cpu-made" if same else "differs from cpu-made: max |d| = %.3g on %d of %d"
                            % (np.abs(cpu[k].astype(np.float64) - out[k]).max(), (cpu[k] != out[k]).sum(), out[k].size)))
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    if "--gpu" in sys.argv:
        return main_gpu()
    out = run_all(load_reference())
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    path = os.path.join(ROOT, "tests", "golden", "pointnet2_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
