"""Generates tests/golden/interp.npz by running the REFERENCE's own three_nn / three_interpolate(+grad) CUDA ops
(extensions/pointnet2/_ext_src, rebuilt unmodified for sm_100a by oracle/build_ref.py -> oracle/_ref/) on a B200:

    gpurun -- 'python tests/golden/make_golden_interp.py gpurun_out/golden'   # then copy into tests/golden/

Inputs are stored next to the outputs.  Same role as make_golden.py: the pins of the oracle and of the kernels.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pointdae_b200 import synth  # noqa: E402
import _interp_cases  # noqa: E402
import _refmods  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    ext = _refmods.ref_pointnet2()
    assert ext is not None, "oracle/_ref not built (python oracle/build_ref.py)"
    blob = {}
    for name, (unknown, known) in _interp_cases.nn_cases(synth).items():
        u, k = torch.from_numpy(unknown).to(dev), torch.from_numpy(known).to(dev)
        dist2, idx = ext.three_nn(u, k)
        d2, ix = dist2.cpu().numpy(), idx.cpu().numpy()
        blob[name + "/unknown"], blob[name + "/known"] = unknown, known
        blob[name + "/dist2"], blob[name + "/idx"] = d2, ix
        for c in (5, 16):
            feats, weight, gout = _interp_cases.interp_inputs(synth, name, unknown, known, ix, d2, c)
            f, w, g = (torch.from_numpy(a).to(dev) for a in (feats, weight, gout))
            out = ext.three_interpolate(f, idx, w)
            gp = ext.three_interpolate_grad(g, idx, w, known.shape[1])
            torch.cuda.synchronize()
            blob["%s/c%d/feats" % (name, c)], blob["%s/c%d/weight" % (name, c)] = feats, weight
            blob["%s/c%d/gout" % (name, c)] = gout
            blob["%s/c%d/out" % (name, c)] = out.cpu().numpy()
            blob["%s/c%d/gfeats" % (name, c)] = gp.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "interp.npz"), **blob)
    print("golden written:", os.path.join(out_dir, "interp.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
