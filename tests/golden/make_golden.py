"""Generates tests/golden/*.npz by running the REFERENCE's own CUDA ops (rebuilt unmodified for
sm_100a by oracle/build_ref.py -> oracle/_ref/) on a B200:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy into tests/golden/

These are the pins of the CPU oracle and of the new kernels: the reference's own test-suite holds
no golden vector for any op on this path, so the pins are outputs of the
reference itself.  Inputs are stored next to the outputs so no RNG needs to be reproduced.
KNN_CUDA is not vendored in the reference tree, hence there is no golden file for kNN / Group
(those are pinned against the oracle only -- "parity unpinned").
The DGCNN knn / get_graph_feature fixture is produced from a verbatim evaluation of the reference's
pure-torch formula on the GPU (models/dgcnn_util.py:7-36 restated below, as /root/reference is not
on the GPU box).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pointdae_b200 import synth  # noqa: E402
import _refmods  # noqa: E402


def fps_cases():
    cases = {}
    cases["c2_1024_64"] = (synth.clouds(4, 1024, seed=11), 64)
    cases["h_2048_64"] = (synth.clouds(2, 2048, seed=12), 64)
    cases["adv_1024_128"] = (synth.adversarial(synth.clouds(3, 1024, seed=13), seed=13), 128)
    cases["n1000_96"] = (synth.adversarial(synth.clouds(2, 1000, seed=14), seed=14), 96)
    cases["n100_40"] = (synth.clouds(2, 100, seed=15), 40)
    cases["full_256_256"] = (synth.adversarial(synth.clouds(2, 256, seed=16), seed=16, n_small=4, n_dup=8), 256)
    z = synth.clouds(2, 300, seed=17)
    z[0] = 0.0  # an all-zero cloud: every point skipped
    cases["zero_300_16"] = (z, 16)
    cases["n8192_512"] = (synth.clouds(1, 8192, seed=18), 512)
    cases["n5000_64"] = (synth.adversarial(synth.clouds(1, 5000, seed=19), seed=19), 64)
    return cases


def chamfer_cases():
    cases = {}
    a = synth.clouds(4, 1024, seed=21)
    cases["c3_1024"] = (synth.prediction(a, seed=21), a)
    a = synth.clouds(2, 2048, seed=22)
    cases["h_2048"] = (synth.prediction(a, seed=22), a)
    a = synth.clouds(2, 1300, seed=23)
    cases["ragged_700_1300"] = (synth.clouds(2, 700, seed=24), a)
    a = synth.adversarial(synth.clouds(2, 600, seed=25), seed=25, n_small=0, n_dup=64)
    cases["ties_600"] = (a[:, ::-1].copy(), a)
    t = synth.clouds(64, 36, seed=26)
    cases["tiny_36_32"] = (t, t[:, :32] + np.float32(0.01))
    cases["coarse_64_64"] = (synth.clouds(16, 64, seed=27), synth.clouds(16, 64, seed=28))
    cases["one_1_5"] = (synth.clouds(3, 1, seed=29), synth.clouds(3, 5, seed=30))
    cases["mid_200_130"] = (synth.clouds(3, 200, seed=31), synth.clouds(3, 130, seed=32))
    return cases


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    ext = _refmods.ref_pointnet2()
    cham = _refmods.ref_chamfer()
    assert ext is not None and cham is not None, "oracle/_ref not built (python oracle/build_ref.py)"

    # ---- FPS + gather (+grad) --------------------------------------------------------------------
    blob = {}
    for name, (xyz, m) in fps_cases().items():
        t = torch.from_numpy(xyz).to(dev)
        idx = ext.furthest_point_sampling(t, m)
        feat = t.transpose(1, 2).contiguous()
        g = ext.gather_points(feat, idx)
        blob[name + "/xyz"] = xyz
        blob[name + "/npoint"] = np.int32(m)
        blob[name + "/idx"] = idx.cpu().numpy()
        blob[name + "/gathered"] = g.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "fps_gather.npz"), **blob)

    # ---- Chamfer fwd/bwd ------------------------------------------------------------------------
    blob = {}
    for name, (x1, x2) in chamfer_cases().items():
        t1, t2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev)
        d1, d2, i1, i2 = cham.forward(t1, t2)
        rng = np.random.default_rng(5)
        g1 = rng.uniform(0.5, 1.5, size=d1.shape).astype(np.float32) / d1.numel()
        g2 = rng.uniform(0.5, 1.5, size=d2.shape).astype(np.float32) / d2.numel()
        gx1, gx2 = cham.backward(t1, t2, i1, i2, torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev))
        torch.cuda.synchronize()
        for k, v in (("xyz1", x1), ("xyz2", x2), ("dist1", d1), ("dist2", d2), ("idx1", i1), ("idx2", i2),
                     ("gd1", g1), ("gd2", g2), ("gx1", gx1), ("gx2", gx2)):
            blob[name + "/" + k] = v.cpu().numpy() if torch.is_tensor(v) else v
    # the transposed, non-contiguous input of models/PointCAE_transformer.py:1059-1066
    base = torch.from_numpy(synth.clouds(8, 36, seed=33)).to(dev)  # (8,36,3) values
    conv_out = base.transpose(1, 2).contiguous()  # (8,3,36) "folding" output
    view = conv_out.transpose(1, 2)  # (8,36,3) non-contiguous view, passed straight in
    tgt = torch.from_numpy(synth.clouds(8, 32, seed=34)).to(dev)
    d1, d2, i1, i2 = cham.forward(view, tgt)
    gd1 = torch.full_like(d1, 1.0 / d1.numel())
    gd2 = torch.full_like(d2, 1.0 / d2.numel())
    gx1, gx2 = cham.backward(view, tgt, i1, i2, gd1, gd2)
    torch.cuda.synchronize()
    blob["transposed/conv_out"] = conv_out.cpu().numpy()
    blob["transposed/xyz2"] = tgt.cpu().numpy()
    for k, v in (("dist1", d1), ("dist2", d2), ("idx1", i1), ("idx2", i2), ("gx2", gx2)):
        blob["transposed/" + k] = v.cpu().numpy()
    blob["transposed/gx1_strides"] = np.array(gx1.stride(), dtype=np.int64)
    # raw storage order of gx1: it is a transposed view of a (8,3,36)-ordered buffer iff strides were preserved
    blob["transposed/gx1_storage"] = gx1.transpose(1, 2).contiguous().cpu().numpy() if gx1.stride() == view.stride() \
        else gx1.contiguous().cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "chamfer.npz"), **blob)

    # ---- ball query + grouping ----------------------------------------------------------------------
    blob = {}
    xyz = synth.clouds(2, 4096, seed=41)
    t = torch.from_numpy(xyz).to(dev)
    cidx = ext.furthest_point_sampling(t, 256)
    new_xyz = ext.gather_points(t.transpose(1, 2).contiguous(), cidx).transpose(1, 2).contiguous()
    for radius, ns in ((0.2, 64), (0.05, 16), (0.4, 8)):
        bq = ext.ball_query(new_xyz, t, radius, ns)
        feats = t.transpose(1, 2).contiguous()
        gp = ext.group_points(feats, bq)
        key = "r%g_s%d" % (radius, ns)
        blob[key + "/idx"] = bq.cpu().numpy()
        blob[key + "/grouped_sum"] = gp.sum(dim=(2, 3)).cpu().numpy()  # checksum of the grouped tensor
        if ns == 16:
            blob[key + "/grouped"] = gp.cpu().numpy()
    blob["xyz"] = xyz
    blob["new_xyz"] = new_xyz.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "ball_group.npz"), **blob)

    # ---- DGCNN knn / get_graph_feature: the reference's torch formula, evaluated verbatim ------------
    blob = {}
    for name, (c, n, k) in {"c3": (3, 512, 20), "c64": (64, 256, 20)}.items():
        x = torch.from_numpy(synth.features(2, c, n, seed=51 + c)).to(dev)
        inner = -2 * torch.matmul(x.transpose(2, 1), x)  # models/dgcnn_util.py:8
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        pd = -xx - inner - xx.transpose(2, 1)
        idx = pd.topk(k=k, dim=-1)[1]
        xt = x.transpose(2, 1).contiguous()
        feat = torch.gather(xt.unsqueeze(1).expand(-1, n, -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, c))
        full = torch.cat((feat - xt.unsqueeze(2), xt.unsqueeze(2).expand(-1, -1, k, -1)), dim=3).permute(0, 3, 1, 2)
        blob[name + "/x"] = x.cpu().numpy()
        blob[name + "/idx"] = idx.cpu().numpy()
        blob[name + "/feature_sum_k"] = full.sum(dim=3).cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "dgcnn.npz"), **blob)
    print("golden written to", out_dir, os.listdir(out_dir))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
