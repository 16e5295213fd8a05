"""Generates tests/golden/edgeconv_ref.npz in THIS container (CPU only):

    python tests/golden/make_golden_edgeconv.py

Runs the four EdgeConv layers of the reference's OWN `dgcnn_encoder` (models/dgcnn_util.py:96-128, imported from
/root/reference, eval mode, seeded weights and non-trivial BatchNorm statistics) layer by layer and stores, per layer,
the input, the neighbour indices its own `knn` picked, the convolution weight, the folded BatchNorm scale / shift and the
output `max over k`.  Pins the oracle of SURVEY.md 8f row 4 (kernel to come)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointdae_b200 import synth  # noqa: E402

REF = "/root/reference/models/dgcnn_util.py"


def main():
    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    spec = importlib.util.spec_from_file_location("ref_dgcnn_util", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(11)
    enc = ref.dgcnn_encoder(channel=3).eval()
    with torch.no_grad():
        for bn in (enc.bn1, enc.bn2, enc.bn3, enc.bn4):  # trained-looking statistics, some negative scales
            bn.weight.copy_(torch.randn_like(bn.weight))
            bn.bias.copy_(0.3 * torch.randn_like(bn.bias))
            bn.running_mean.copy_(0.2 * torch.randn_like(bn.running_mean))
            bn.running_var.copy_(0.5 + torch.rand_like(bn.running_var))
    x = torch.from_numpy(synth.features(1, 3, 192, seed=8))
    out = {}
    with torch.no_grad():
        for li, (conv, bn) in enumerate(((enc.conv1, enc.bn1), (enc.conv2, enc.bn2), (enc.conv3, enc.bn3), (enc.conv4, enc.bn4))):
            k = 20
            idx = ref.knn(x, k)
            feat = ref.get_graph_feature(x, k=k, idx=idx.clone())
            y = conv(feat).max(dim=-1, keepdim=False)[0]
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            out["l%d/x" % li], out["l%d/idx" % li] = x.numpy(), idx.numpy()
            out["l%d/weight" % li] = conv[0].weight.view(conv[0].weight.size(0), -1).numpy()
            out["l%d/scale" % li], out["l%d/shift" % li] = scale.numpy(), (bn.bias - scale * bn.running_mean).numpy()
            out["l%d/out" % li] = y.numpy()
            print("layer", li, tuple(x.shape), "->", tuple(y.shape))
            x = y
    path = os.path.join(ROOT, "tests", "golden", "edgeconv_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
